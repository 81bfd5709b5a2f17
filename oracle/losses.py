"""torch restatement of the NeRF training losses -- TEST INFRASTRUCTURE ONLY.

Follows nr4seg/lightning/joint_train_lightning_net.py:199-221 (semantic zero-mass fix-up, colour MSE,
NLL(log(p + 1e-15), ignore -1), depth L1 over gt_depth != 0 divided by one_m_to_scene_uom) and :503-507
(total = colour + 0.04 semantics + 0.1 depth).  The product path computes the same numbers and their gradients in
one kernel (csrc/loss.cu, ucsa_nerf_loss); tests hold that kernel to this expression.
"""
from __future__ import annotations

import torch


def nerf_losses(outputs, gt_rgb, labels, gt_depth, one_m_to_scene_uom, weight_depth=0.1, weight_semantics=0.04,
                global_scale=1.0):
    """Shapes [B,N,...] as rendered.  -> (total * global_scale, (colour, semantic, depth))"""
    pred_rgb, semantics, pred_depth = outputs["image"], outputs["semantics"], outputs["depth"]
    labels = labels.clone()
    invalid = torch.sum(semantics, dim=-1) == 0
    semantics = torch.where(invalid.unsqueeze(-1), torch.ones_like(semantics), semantics)
    semantics = semantics / torch.sum(semantics, dim=-1, keepdim=True)
    labels[invalid] = -1
    loss_color = torch.nn.functional.mse_loss(pred_rgb, gt_rgb.float(), reduction="none").mean()
    logp = torch.log(semantics + 1e-15).permute(0, 2, 1)
    loss_sem = torch.nn.functional.nll_loss(logp, labels, ignore_index=-1, reduction="none").mean()
    valid = gt_depth != 0
    loss_depth = torch.nn.functional.l1_loss(pred_depth[valid] / one_m_to_scene_uom, gt_depth[valid],
                                             reduction="none").mean(-1)
    total = loss_color + loss_sem * weight_semantics + loss_depth * weight_depth
    return total * global_scale, (loss_color, loss_sem, loss_depth)
