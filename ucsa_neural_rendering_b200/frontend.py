"""Device-resident views -> ray batches (row f2 of SURVEY.md section 8).

The reference builds every training batch with eager torch ops on the whole pixel grid of a view
(JointTrainLightningNet.get_rays_train + forward_nerf_train, joint_train_lightning_net.py:109-151,167-187): two H*W
mesh grids, a gather, a normalisation and a 3x3 product, then three more gathers for the ground truth.  Here the views
(fp16 image planes, labels, depth, poses) stay on the device and a batch is two kernel launches
(ucsa_generate_rays, ucsa_gather_gt) on N sampled pixels."""
from __future__ import annotations

import torch

from . import ops


class ViewSampler:

    def __init__(self, poses, intrinsics, height, width, images_h=None, labels=None, depths=None, device="cuda"):
        """poses [V,4,4] float32 cam2world (instant-ngp convention, poses.nerf_matrix_to_ngp); intrinsics
        (fx, fy, cx, cy); images_h [V,C,H,W] fp16 (batch["img_fp16"]); labels [V,H,W] int64; depths [V,H,W] f32."""
        self.device = torch.device(device)
        self.height, self.width = int(height), int(width)
        self.intrinsics = tuple(float(v) for v in intrinsics)
        self.poses = torch.as_tensor(poses, dtype=torch.float32).to(self.device).contiguous()
        self.images_h = None if images_h is None else torch.as_tensor(images_h).to(self.device, torch.float16).contiguous()
        self.labels = None if labels is None else torch.as_tensor(labels).to(self.device, torch.int64).contiguous()
        self.depths = None if depths is None else torch.as_tensor(depths).to(self.device, torch.float32).contiguous()

    @property
    def n_views(self):
        return self.poses.shape[0]

    def sample(self, view: int, n: int = 4096, generator=None):
        """N pixels of one view drawn with replacement (torch.randint, as joint_train_lightning_net.py:142).
        -> dict(rays_o, rays_d [N,3], direction_norms [N], inds [N], gt_rgb fp16 [N,C], labels [N], depth [N])"""
        n = min(n, self.height * self.width)
        inds = torch.randint(0, self.height * self.width, (n,), device=self.device, generator=generator)
        rays_o, rays_d, norms = ops.generate_rays(self.poses[view], self.intrinsics, self.height, self.width, inds=inds)
        out = {"rays_o": rays_o, "rays_d": rays_d, "direction_norms": norms, "inds": inds}
        if self.images_h is not None:
            gt_rgb, gt_labels, gt_depth = ops.gather_gt(
                self.images_h[view], inds, labels=None if self.labels is None else self.labels[view],
                depth=None if self.depths is None else self.depths[view])
            out.update(gt_rgb=gt_rgb, labels=gt_labels, depth=gt_depth)
        return out

    def full_view(self, view: int):
        """all H*W rays of a view in row-major pixel order (ngp_utils.get_rays)"""
        return ops.generate_rays(self.poses[view], self.intrinsics, self.height, self.width)
