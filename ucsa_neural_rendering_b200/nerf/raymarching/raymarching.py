"""Mirror of nr4seg/nerf/raymarching/raymarching.py on libucsa_nerf.so (no JIT build at import time).

Same callables, argument order and defaults as the reference's autograd wrappers; inputs are cast to fp32
like ``custom_fwd(cast_inputs=torch.float32)`` does there."""
import torch

from ... import ops


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:12-47.  rays_o, rays_d: [N,3]; aabb: [6] -> nears, fars: [N]."""
    if not rays_o.is_cuda:
        rays_o = rays_o.cuda()
    if not rays_d.is_cuda:
        rays_d = rays_d.cuda()
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    return ops.near_far_from_aabb(rays_o, rays_d, aabb.float().contiguous().to(rays_o.device), min_near)


__all__ = ["near_far_from_aabb"]
