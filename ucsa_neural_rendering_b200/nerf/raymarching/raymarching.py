"""Mirror of nr4seg/nerf/raymarching/raymarching.py on libucsa_nerf.so (no JIT build at import time).

Same callables, argument order and defaults as the reference's autograd wrappers (raymarching.py:12-595); inputs are
cast to fp32 like ``custom_fwd(cast_inputs=torch.float32)`` does there.  The three wrappers whose backend symbols are
missing in the reference (``composite_rays_train_semantics``, ``composite_rays``, ``composite_rays_semantics``; they
raise AttributeError there) are functional here.  An optional ``bitfield=`` keyword lets the marchers read the packed
occupancy bits instead of the float grid (identical samples, far less traffic)."""
import torch
from torch.autograd import Function

from ... import ops


def _f32(t):
    return t.float().contiguous()


# ----------------------------------------------------------------------------------------------- utils
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:12-47.  rays_o, rays_d: [N,3]; aabb: [6] -> nears, fars: [N]."""
    if not rays_o.is_cuda:
        rays_o = rays_o.cuda()
    if not rays_d.is_cuda:
        rays_d = rays_d.cuda()
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    return ops.near_far_from_aabb(rays_o, rays_d, _f32(aabb).to(rays_o.device), min_near)


# ----------------------------------------------------------------------------------------------- train
def march_rays_train(rays_o, rays_d, bound, density_grid, mean_density, nears, fars, step_counter=None, mean_count=-1,
                     perturb=False, align=-1, force_all_rays=False, dt_gamma=0, bitfield=None):
    """raymarching.py:54-166 -> xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3] (ray id, offset, count)."""
    if not rays_o.is_cuda:
        rays_o = rays_o.cuda()
    if not rays_d.is_cuda:
        rays_d = rays_d.cuda()
    if not density_grid.is_cuda:
        density_grid = density_grid.cuda()
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    density_grid = _f32(density_grid)
    n = rays_o.shape[0]
    m = n * 1024  # raymarching.py:109
    if not force_all_rays and mean_count > 0:
        if align > 0:
            mean_count += align - mean_count % align
        m = mean_count
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=rays_o.device)
    xyzs, dirs, deltas, rays = ops.march_rays_train(rays_o, rays_d, density_grid, bitfield, mean_density, bound,
                                                    dt_gamma, _f32(nears), _f32(fars), m, step_counter, perturb)
    if force_all_rays or mean_count <= 0:
        m = step_counter[0].item()  # D2H copy, as in the reference (raymarching.py:153-161)
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    return xyzs, dirs, deltas, rays


class _composite_rays_train(Function):
    """raymarching.py:169-246 (rgb + depth; depth gets no gradient, :209)."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays):
        sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
        n = rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(n, dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        image = torch.empty(n, 3, dtype=torch.float32, device=dev)
        ops.composite_rays_train_forward(sigmas, rgbs, None, deltas, rays, 0, weights_sum, depth, image, None)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image)
        return weights_sum, depth, image

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        sigmas, rgbs, deltas, rays, weights_sum, image = ctx.saved_tensors
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        ops.composite_rays_train_backward(_f32(grad_weights_sum), _f32(grad_image), None, sigmas, rgbs, deltas, rays,
                                          weights_sum, image, 0, grad_sigmas, grad_rgbs, None)
        return grad_sigmas, grad_rgbs, None, None


composite_rays_train = _composite_rays_train.apply


class _composite_rays_train_semantics(Function):
    """raymarching.py:249-360; the backend kernels are commented out in the reference (raymarching.h:12-13)."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, local_semantics, deltas, rays, num_semantics_classes):
        sigmas, rgbs, local_semantics, deltas = _f32(sigmas), _f32(rgbs), _f32(local_semantics), _f32(deltas)
        n = rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(n, dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        image = torch.empty(n, 3, dtype=torch.float32, device=dev)
        semantics = torch.empty(n, num_semantics_classes, dtype=torch.float32, device=dev)
        ops.composite_rays_train_forward(sigmas, rgbs, local_semantics, deltas, rays, num_semantics_classes, weights_sum,
                                         depth, image, semantics)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image)
        ctx.n_classes = num_semantics_classes
        return weights_sum, depth, image, semantics

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_depth, grad_image, grad_semantics):
        sigmas, rgbs, deltas, rays, weights_sum, image = ctx.saved_tensors
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        grad_local = torch.zeros(sigmas.shape[0], ctx.n_classes, dtype=torch.float32, device=sigmas.device)
        ops.composite_rays_train_backward(_f32(grad_weights_sum), _f32(grad_image), _f32(grad_semantics), sigmas, rgbs,
                                          deltas, rays, weights_sum, image, ctx.n_classes, grad_sigmas, grad_rgbs,
                                          grad_local)
        return grad_sigmas, grad_rgbs, grad_local, None, None, None


composite_rays_train_semantics = _composite_rays_train_semantics.apply


# ----------------------------------------------------------------------------------------------- infer
def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_grid, mean_density, near, far,
               align=-1, perturb=False, dt_gamma=0, bitfield=None):
    """raymarching.py:367-452 -> xyzs, dirs, deltas of n_alive * n_step (padded to `align`) samples."""
    if not rays_o.is_cuda:
        rays_o = rays_o.cuda()
    if not rays_d.is_cuda:
        rays_d = rays_d.cuda()
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    m = n_alive * n_step
    if align > 0:
        m += align - (m % align)
    return ops.march_rays(n_alive, n_step, rays_alive, _f32(rays_t), rays_o, rays_d, bound, dt_gamma,
                          _f32(density_grid), bitfield, mean_density, _f32(near), _f32(far), m, int(perturb))


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image,
                   num_semantics_classes=None):
    """raymarching.py:455-504 (in-place accumulation into weights_sum / depth / image, rays_t updated)."""
    ops.composite_rays(n_alive, n_step, rays_alive, rays_t, _f32(sigmas), _f32(rgbs), None, _f32(deltas), 0,
                       weights_sum, depth, image, None)
    return tuple()


def composite_rays_semantics(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, local_semantics, deltas, weights_sum,
                             depth, image, semantics, logits=None):
    """raymarching.py:507-558; as composite_rays plus the in-place [N,C] semantic accumulator.  `logits` (fp16
    [M, >= C], contiguous) may replace `local_semantics`: the soft-max is then taken inside the kernel."""
    ops.composite_rays(n_alive, n_step, rays_alive, rays_t, _f32(sigmas), _f32(rgbs),
                       None if logits is not None else _f32(local_semantics), _f32(deltas), semantics.shape[-1],
                       weights_sum, depth, image, semantics, logits=logits)
    return tuple()


def compact_rays(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter):
    """raymarching.py:561-595 (order preserving here)."""
    ops.compact_rays(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter)
    return tuple()


__all__ = ["near_far_from_aabb", "march_rays_train", "composite_rays_train", "composite_rays_train_semantics",
           "march_rays", "composite_rays", "composite_rays_semantics", "compact_rays"]
