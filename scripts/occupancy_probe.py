#!/usr/bin/env python
"""The occupancy-grid path alone (rows a16-a20): grid refresh, one training render (forward + backward) and one
inference view, for `ncu --metrics gpu__time_duration.sum` launch lists.  python scripts/occupancy_probe.py"""
import sys

import torch

sys.path.insert(0, ".")
from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork  # noqa: E402
from ucsa_neural_rendering_b200.scene import SyntheticScene  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    scene = SyntheticScene(seed=0, device=dev)
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=True, density_scale=1, num_semantic_classes=40).to(dev)
    with torch.no_grad():
        net.encoder.params.uniform_(-0.5, 0.5)
    net.train()
    g = torch.Generator(device=dev).manual_seed(1)
    pix = torch.randint(0, scene.W * scene.H, (4096,), device=dev, generator=g)
    o, d, dn = scene.rays(0, pix)
    for it in range(2):
        net.update_extra_state()
        out = net.render(o[None], d[None], direction_norms=dn.view(1, -1, 1), staged=False, perturb=True, dt_gamma=1 / 128,
                         force_all_rays=True)
        (out["image"].sum() + out["semantics"].sum()).backward()
    torch.cuda.synchronize()
    print("samples", int(net.step_counter[(net.local_step - 1) % 16, 0]), "mean density", net.mean_density)
    net.eval()
    with torch.no_grad():
        vo, vd, vdn = scene.rays(0, torch.arange(scene.W * scene.H, device=dev))
        out = net.render(vo[None], vd[None], direction_norms=vdn.view(1, -1, 1), staged=True, perturb=False, dt_gamma=1 / 128)
    torch.cuda.synchronize()
    print("ok", float(out["image"].mean()))
