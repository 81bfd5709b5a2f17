#!/usr/bin/env python
"""The occupancy-grid path alone (rows a16-a20): grid refresh, the training render (forward + backward) with the fused
node and with the module-level heads, and one inference view for a sweep of wavefront schedules.  Also the target of
`ncu --metrics gpu__time_duration.sum` launch lists (profiles/r2_occupancy_*.md).

    python scripts/occupancy_probe.py [--train-only]"""
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
    def train_render():
        net.zero_grad(set_to_none=True)
        out = net.render(o[None], d[None], direction_norms=dn.view(1, -1, 1), staged=False, perturb=True, dt_gamma=1 / 128,
                         force_all_rays=True)
        (out["image"].sum() + out["semantics"].sum()).backward()

    for fused in (True, False):
        net.fused_packed = fused
        net.update_extra_state()
        train_render()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            train_render()
        e1.record()
        torch.cuda.synchronize()
        print("training render forward + backward,", "fused heads node" if fused else "module-level heads",
              f"{e0.elapsed_time(e1) / 5:.3f} ms")
    net.fused_packed = True
    print("samples", int(net.step_counter[(net.local_step - 1) % 16, 0]), "mean density", net.mean_density)
    if "--train-only" in sys.argv:
        sys.exit(0)
    net.eval()
    with torch.no_grad():
        vo, vd, vdn = scene.rays(0, torch.arange(scene.W * scene.H, device=dev))
        ref = None
        for steps in ((1, 8), (2, 8), (4, 8), (8, 8), (8, 16), (16, 16), (16, 32), (32, 32)):
            net.wavefront_steps = steps
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = net.render(vo[None], vd[None], direction_norms=vdn.view(1, -1, 1), staged=True, perturb=False,
                                 dt_gamma=1 / 128)
                e1.record()
                torch.cuda.synchronize()
            if ref is None:
                ref = out
            diff = max(float((out[k] - ref[k]).abs().max()) for k in ("image", "depth", "semantics"))
            print(f"inference view, steps per round {steps}: {e0.elapsed_time(e1):.1f} ms, max |diff| to (1, 8): {diff:.2e}")
    print("ok", float(out["image"].mean()))
