"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the two ends of the rendering path.

* get_rays            - nr4seg/dataset/ngp_utils.py:28-70 (and get_rays_train, joint_train_lightning_net.py:109-151,
                        which is the same arithmetic on the gathered pixel indices)
* gather_gt           - the three torch.gather of forward_nerf_train (joint_train_lightning_net.py:180-187)
* label_epilogue      - forward_nerf_test / the predict step (joint_train_lightning_net.py:246-250, :755-768)
Pinned by tests/golden/frontend.npz, which tests/golden/make_golden.py generates by running the reference's own
ngp_utils.get_rays / nerf_matrix_to_ngp and replaying the quoted lines."""
from __future__ import annotations

import numpy as np
import torch


def get_rays(pose, intrinsics, height, width, inds=None):
    """pose [4,4] f32 cam2world (ngp convention); -> rays_o [N,3], rays_d [N,3], direction_norms [N]"""
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    pose = torch.as_tensor(pose, dtype=torch.float32)
    if inds is None:
        inds = torch.arange(height * width)
    inds = torch.as_tensor(inds, dtype=torch.int64)
    i = (inds % width).float() + 0.5
    j = torch.div(inds, width, rounding_mode="floor").float() + 0.5
    zs = torch.ones_like(i)
    xs = (i - cx) / fx * zs
    ys = (j - cy) / fy * zs
    directions = torch.stack((xs, ys, zs), dim=-1)
    norms = torch.norm(directions, dim=-1, keepdim=True)
    directions = directions / norms
    rays_d = directions @ pose[:3, :3].transpose(-1, -2)
    rays_o = pose[:3, 3].expand_as(rays_d)
    return rays_o, rays_d, norms[:, 0]


def gather_gt(image_chw, labels_hw, depth_hw, inds):
    inds = torch.as_tensor(inds, dtype=torch.int64)
    c = image_chw.shape[0]
    gt_rgb = torch.as_tensor(image_chw).reshape(c, -1).permute(1, 0)[inds]
    return gt_rgb, torch.as_tensor(labels_hw).reshape(-1)[inds], torch.as_tensor(depth_hw).reshape(-1)[inds]


def label_epilogue(semantics, rgb):
    sem = torch.as_tensor(semantics, dtype=torch.float32).clone()
    invalid = torch.sum(sem, dim=-1) == 0
    sem[invalid] = 1
    sem = sem / torch.sum(sem, dim=-1, keepdim=True)
    label_u8 = (torch.argmax(sem, dim=-1) + 1).numpy().astype(np.uint8)
    rgb_u8 = (np.asarray(rgb, dtype=np.float32) * 255).astype(np.uint8)
    return label_u8, rgb_u8
