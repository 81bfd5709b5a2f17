"""Synthetic ScanNet-shaped scenes: an axis-aligned room with cuboid furniture,
hand-held-like camera loop, analytic colour / depth / class ground truth.  Workload generator for bench.py and
the tests; set-up code, not part of the rendering path (plain torch ops, runs once before the timed region).

Conventions follow the reference's data path: pinhole rays with +0.5 pixel centres, unit directions and
``direction_norms`` as in joint_train_lightning_net.py:108-157 / dataset/ngp_utils.py:28-69; cameras at a mean
radius of 0.33*bound like preprocessing_scripts/scannet2nerf.py:187; depth in metres through
``one_m_to_scene_uom``."""
from __future__ import annotations

import math

import torch


class SyntheticScene:

    def __init__(self, seed: int = 0, width: int = 640, height: int = 480, n_views: int = 400, n_classes: int = 40,
                 bound: float = 4.0, device="cpu"):
        g = torch.Generator().manual_seed(1000 + seed)
        self.W, self.H, self.n_views, self.n_classes, self.bound = width, height, n_views, n_classes, bound
        self.device = torch.device(device)
        self.one_m_to_scene_uom = 0.6  # scene units per metre
        r = lambda *s: torch.rand(*s, generator=g)
        half = torch.tensor([2.4, 2.0, 1.1]) + 0.6 * r(3)  # room half extents (scene units), inside bound=4
        boxes_lo = [-half]
        boxes_hi = [half]
        n_furn = 8 + int(r(1).item() * 5)
        for _ in range(n_furn):
            size = 0.15 + 0.5 * r(3)
            cx = (r(1).item() * 2 - 1) * (half[0] - size[0] - 0.05)
            cy = (r(1).item() * 2 - 1) * (half[1] - size[1] - 0.05)
            # keep the camera loop (radius ~1.32, height 0) mostly unobstructed: furniture sits on the floor
            lo = torch.tensor([cx, cy, -half[2].item()]) - torch.tensor([size[0], size[1], 0.0])
            hi = torch.tensor([cx, cy, -half[2].item()]) + torch.tensor([size[0], size[1], 2 * size[2].item()])
            boxes_lo.append(lo)
            boxes_hi.append(hi)
        self.boxes_lo = torch.stack(boxes_lo).to(self.device)  # [B,3]; box 0 is the room (seen from inside)
        self.boxes_hi = torch.stack(boxes_hi).to(self.device)
        nb = self.boxes_lo.shape[0]
        # per box, per face (6) colour and class; the room's faces are wall / floor / ceiling classes
        self.face_rgb = (0.15 + 0.8 * r(nb, 6, 3)).to(self.device)
        cls = torch.randint(3, n_classes, (nb, 1), generator=g).expand(nb, 6).clone()
        cls[0] = torch.tensor([0, 0, 0, 0, 1, 2])  # walls, walls, floor, ceiling
        self.face_cls = cls.to(self.device)

        # camera loop: radius 0.33*bound on average, looking roughly inwards / around, small height wobble
        radius = 0.33 * bound
        ang = torch.linspace(0, 2 * math.pi, n_views + 1)[:-1]
        rad = radius * (0.85 + 0.3 * torch.sin(3 * ang))
        rad = torch.minimum(rad, torch.full_like(rad, float(min(half[0], half[1])) - 0.3))
        pos = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang), 0.15 * torch.sin(5 * ang)], dim=1)
        look = torch.stack([torch.cos(ang + 2.2), torch.sin(ang + 2.2), -0.15 + 0.1 * torch.sin(2 * ang)], dim=1)
        fwd = torch.nn.functional.normalize(look, dim=1)
        up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
        right = torch.nn.functional.normalize(torch.cross(fwd, up, dim=1), dim=1)
        down = torch.cross(fwd, right, dim=1)
        pose = torch.eye(4).repeat(n_views, 1, 1)
        pose[:, :3, 0], pose[:, :3, 1], pose[:, :3, 2], pose[:, :3, 3] = right, down, fwd, pos
        self.poses = pose.to(self.device)  # cam2world, camera looks along +z, y down
        f = 289.0 * (width / 320.0)
        self.intrinsics = (f, f, width / 2.0, height / 2.0)

    # joint_train_lightning_net.py:108-157 for given pixel ids of one view
    def rays(self, view: int, pix: torch.Tensor):
        fx, fy, cx, cy = self.intrinsics
        i = (pix % self.W).float() + 0.5
        j = torch.div(pix, self.W, rounding_mode="floor").float() + 0.5
        dirs = torch.stack([(i - cx) / fx, (j - cy) / fy, torch.ones_like(i)], dim=-1)
        dn = torch.norm(dirs, dim=-1, keepdim=True)
        dirs = dirs / dn
        pose = self.poses[view]
        rays_d = dirs @ pose[:3, :3].T
        rays_o = pose[:3, 3].expand_as(rays_d)
        return rays_o.contiguous(), rays_d.contiguous(), dn.contiguous()

    def ground_truth(self, rays_o, rays_d, dn):
        """-> rgb [N,3] in [0,1], z-depth [N] in metres, label [N] int64"""
        inv = 1.0 / rays_d
        t_lo = (self.boxes_lo[None] - rays_o[:, None]) * inv[:, None]  # [N,B,3]
        t_hi = (self.boxes_hi[None] - rays_o[:, None]) * inv[:, None]
        t_in = torch.minimum(t_lo, t_hi)
        t_out = torch.maximum(t_lo, t_hi)
        enter, enter_ax = t_in.max(dim=-1)
        leave, leave_ax = t_out.min(dim=-1)
        hit = (enter < leave) & (leave > 0)
        # furniture: first entry; room (box 0): exit
        t_hit = torch.where(hit & (enter > 0), enter, torch.full_like(enter, float("inf")))
        t_hit[:, 0] = leave[:, 0]
        ax = enter_ax.clone()
        ax[:, 0] = leave_ax[:, 0]
        t_best, b_best = t_hit.min(dim=1)
        ax_best = torch.gather(ax, 1, b_best[:, None]).squeeze(1)
        sign_pos = torch.gather(rays_d, 1, ax_best[:, None]).squeeze(1) > 0
        is_room = b_best == 0
        # face id: axis*2 + (1 if the +axis face); room exits through the face the ray points to
        face = ax_best * 2 + torch.where(is_room, sign_pos, ~sign_pos).long()
        rgb = self.face_rgb[b_best, face]
        p = rays_o + rays_d * t_best[:, None]
        checker = ((torch.floor(p * 2.5).sum(dim=-1)) % 2 == 0).float()[:, None]
        rgb = (rgb * (0.75 + 0.25 * checker)).clamp(0, 1)
        depth_m = t_best / dn.squeeze(-1) / self.one_m_to_scene_uom
        label = self.face_cls[b_best, face]
        return rgb, depth_m, label
