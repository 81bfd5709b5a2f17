"""NeRF training step of the reference (joint_train_lightning_net.py:167-223 losses, :473-513 step,
:897-919 optimizer) around the drop-in network, plus the ray-sharded multi-GPU form of SURVEY.md section 8e.

``nerf_losses`` is the loss arithmetic of ``forward_nerf_train``; ``NerfTrainer.train_step`` is one
``training_step_nerf`` iteration without the Lightning / logging plumbing: render 4096 rays (perturb=True),
losses, backward, Adam (two groups, weight decay on the MLPs only).  The optimizer is the fused
``ucsa_adam_step`` kernel (row f1): one pass per parameter tensor that also refreshes the fp16 working copy.

Multi-GPU: one process per GPU, each rank renders its own slice of the ray batch; after backward one NCCL
all-reduce(sum) over a single flat gradient buffer; every rank then applies the same Adam update, so no
parameter broadcast is needed.  The colour and semantic losses are means over the *global* ray count.  The depth
loss is a mean over valid (gt_depth != 0) rays: each rank normalises by its own valid count and the 1/world factor
averages the per-rank means, which equals the global mean only when the ranks hold equally many valid depths (rays of
a batch are drawn from the same views, so the counts differ by sampling noise only; the single-GPU objective and
the reference's are unchanged)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops, parallel
from ._lib import UcsaError


class _FusedLosses(torch.autograd.Function):
    """The three losses and their gradients in one kernel (ucsa_nerf_loss); backward only scales the stored
    gradients by the incoming scalar (GradScaler's scale, 1/world ...)."""

    @staticmethod
    def forward(ctx, image, depth, semantics, gt_rgb, labels, gt_depth, uom, w_sem, w_depth, global_scale):
        n, c = semantics.shape
        dev = image.device
        loss4 = torch.empty(4, dtype=torch.float32, device=dev)
        grads = torch.empty(n * (4 + c), dtype=torch.float32, device=dev)
        scratch = ops.loss_scratch(dev)  # per call: never shared with another stream's launch
        g_image, g_depth, g_sem = grads[:3 * n].view(n, 3), grads[3 * n:4 * n], grads[4 * n:].view(n, c)
        ops.nerf_loss(image, depth, semantics, gt_rgb, labels, gt_depth, uom, w_sem, w_depth, global_scale, loss4,
                      g_image, g_depth, g_sem, scratch)
        ctx.save_for_backward(grads)
        ctx.n, ctx.c = n, c
        parts = loss4[1:]
        ctx.mark_non_differentiable(parts)
        return loss4[0], parts

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        (grads,) = ctx.saved_tensors
        n, c = ctx.n, ctx.c
        g = grads * g_total
        return g[:3 * n].view(n, 3), g[3 * n:4 * n], g[4 * n:].view(n, c), None, None, None, None, None, None, None


def nerf_losses(outputs, gt_rgb, labels, gt_depth, one_m_to_scene_uom, weight_depth=0.1, weight_semantics=0.04,
                global_scale=1.0):
    """joint_train_lightning_net.py:199-221 and :503-507.  Shapes [B,N,...] as rendered.
    The three losses and their gradients come from one kernel (ucsa_nerf_loss); CUDA tensors only."""
    pred_rgb, semantics, pred_depth = outputs["image"], outputs["semantics"], outputs["depth"]
    c = semantics.shape[-1]
    if not pred_rgb.is_cuda:
        raise UcsaError("nerf_losses needs CUDA tensors (the rendering path has no CPU fallback)")
    if c > ops.MAX_CLASSES:
        raise NotImplementedError(f"nerf_losses: at most {ops.MAX_CLASSES} classes")
    rgb = gt_rgb.reshape(-1, 3)
    if rgb.dtype not in (torch.float16, torch.float32):
        rgb = rgb.float()
    total, parts = _FusedLosses.apply(
        pred_rgb.reshape(-1, 3).float().contiguous(), pred_depth.reshape(-1).float().contiguous(),
        semantics.reshape(-1, c).float().contiguous(), rgb.contiguous(), labels.reshape(-1).long().contiguous(),
        gt_depth.reshape(-1).float().contiguous(), float(one_m_to_scene_uom), float(weight_semantics),
        float(weight_depth), float(global_scale))
    return total, (parts[0], parts[1], parts[2])


class NerfTrainer:

    def __init__(self, net, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay_net=1e-6, num_steps=256,
                 upsample_steps=256, distributed=False):
        self.net = net
        self.lr, self.betas, self.eps, self.wd_net = lr, betas, eps, weight_decay_net
        self.num_steps, self.upsample_steps = num_steps, upsample_steps
        self.distributed = distributed and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.distributed else 1
        self.rank = dist.get_rank() if self.distributed else 0
        self.step = 0
        # parameter tensors in a fixed order; group "encoding" has no weight decay
        self.groups = [(net.encoder, 0.0), (net.sigma_net, weight_decay_net), (net.color_net, weight_decay_net),
                       (net.semantics_net, weight_decay_net)]
        dev = net.encoder.params.device
        sizes = [m.params.numel() for m, _ in self.groups]
        self.flat_grad = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.grad_views, off = [], 0
        for s in sizes:
            self.grad_views.append(self.flat_grad[off:off + s])
            off += s
        self.exp_avg = [torch.zeros_like(m.params) for m, _ in self.groups]
        self.exp_avg_sq = [torch.zeros_like(m.params) for m, _ in self.groups]
        for (m, _), g in zip(self.groups, self.grad_views):
            m.params.grad = g  # autograd accumulates straight into the flat all-reduce buffer
            m.half_params()
        # GradScaler-style overflow handling on the device (joint_train_lightning_net.py:46,509-513)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.skipped_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        self._check_scratch = torch.zeros(2, dtype=torch.int32, device=dev)

    def zero_grad(self):
        self.flat_grad.zero_()

    def optimizer_step(self):
        """Adam on the (already all-reduced) flat gradient; a step whose gradient holds inf / NaN is skipped."""
        self.step += 1
        self.step_dev.add_(1)
        ops.grad_check(self.flat_grad, self.found_inf, self._check_scratch, skipped_dev=self.skipped_dev)
        for (m, wd), g, ea, eas in zip(self.groups, self.grad_views, self.exp_avg, self.exp_avg_sq):
            half = m.half_params()
            ops.adam_step(m.params.data, g, ea, eas, half, lr=self.lr, beta1=self.betas[0], beta2=self.betas[1],
                          eps=self.eps, weight_decay=wd, grad_scale_inv=1.0, found_inf=self.found_inf, step=1,
                          step_dev=self.step_dev, skipped_dev=self.skipped_dev)

    def train_step(self, rays_o, rays_d, direction_norms, gt_rgb, labels, gt_depth, one_m_to_scene_uom, seed=None,
                   ray_base=0):
        """One optimisation step on this rank's rays ([1,n,...] tensors).  Returns the (local) loss tensor."""
        net = self.net
        self.zero_grad()
        out = net.render(rays_o, rays_d, direction_norms=direction_norms, staged=False, bg_color=None, perturb=True,
                         num_steps=self.num_steps, upsample_steps=self.upsample_steps, seed=seed,
                         ray_base=ray_base)
        loss, _ = nerf_losses(out, gt_rgb, labels, gt_depth, one_m_to_scene_uom, global_scale=1.0 / self.world)
        loss.backward()
        for (m, _), g in zip(self.groups, self.grad_views):
            if m.params.grad is not g:  # autograd replaced the tensor: copy back into the flat buffer
                g.copy_(m.params.grad)
                m.params.grad = g
        if self.distributed:
            parallel.all_reduce_gradients(self.flat_grad)
        self.optimizer_step()
        return loss
