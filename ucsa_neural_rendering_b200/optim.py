"""FusedAdam: the fused optimizer kernel behind the torch.optim interface (row f1 of SURVEY.md section 8).

`configure_optimizers` of the reference builds `torch.optim.Adam([...two groups...], lr=1e-2, betas=(0.9, 0.99),
eps=1e-15)` and drives it through `torch.cuda.amp.GradScaler` (joint_train_lightning_net.py:46,509-513,897-919).
`FusedAdam` takes the same arguments and produces the same numbers (ucsa_adam_step follows torch's operation order;
tests hold it to torch.optim.Adam at rtol 1e-6), but

* one kernel per parameter tensor does unscale + inf-skip + moments + update, and also refreshes the fp16 working
  copy the rendering kernels read (torch's foreach Adam is ~10 passes over the 52 MB table plus a separate cast);
* it declares `_step_supports_amp_scaling`, so `GradScaler.step` hands over its scale and its found-inf flag as
  device tensors instead of unscaling the gradients in a separate pass and synchronising the host on the flag.

Drop-in change in the reference: `torch.optim.Adam(...)` -> `FusedAdam(..., network=self.nerf_model)`.
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True  # torch.amp.GradScaler: pass grad_scale / found_inf, do not unscale or sync

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, network=None):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        # (GradScaler sets `self.grad_scale` / `self.found_inf` around each step and deletes them afterwards: they must
        # not exist as attributes in between)
        # parameter -> module that keeps an fp16 working copy of it (refreshed by the same kernel)
        self._half_owner = {}
        if network is not None:
            for mod in network.modules():
                if hasattr(mod, "half_params") and isinstance(getattr(mod, "params", None), torch.nn.Parameter):
                    self._half_owner[id(mod.params)] = mod

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        found_inf, grad_scale = getattr(self, "found_inf", None), getattr(self, "grad_scale", None)
        if found_inf is not None:
            found_inf = found_inf.reshape(1).float()
            step_inc = (found_inf == 0).to(torch.int32)  # 0 on a skipped step
        if grad_scale is not None:
            grad_scale = grad_scale.reshape(1).float()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise ops.UcsaError("FusedAdam: parameters must be contiguous fp32 CUDA tensors")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    # Adam's step count on the device: it must not advance on a step GradScaler skips
                    st["step"] = torch.zeros(1, dtype=torch.int32, device=p.device)
                st["step"].add_(1 if found_inf is None else step_inc)
                owner = self._half_owner.get(id(p))
                half = owner.half_params() if owner is not None else None
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                ops.adam_step(p.data.view(-1), grad.view(-1), st["exp_avg"].view(-1), st["exp_avg_sq"].view(-1),
                              None if half is None else half.view(-1), lr=group["lr"], beta1=b1, beta2=b2,
                              eps=group["eps"], weight_decay=group["weight_decay"], grad_scale_inv=1.0,
                              grad_scale_dev=grad_scale, found_inf=found_inf, step=1, step_dev=st["step"])
                # (the kernel wrote the fp16 copy of the NEW values through raw pointers: the parameter's version is
                # unchanged, so the owner's cached copy stays the valid one)
        return loss
