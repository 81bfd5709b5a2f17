"""trunc_exp of nr4seg/nerf/activation.py:7-19: exp forward in fp32, backward g * exp(clamp(x, -15, 15)).

The fused kernels apply the same rule inside ucsa_density_fwd / ucsa_density_bwd; this autograd function is
the stand-alone export the reference module offers."""
import torch
from torch.autograd import Function


class _trunc_exp(Function):

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply
