#!/usr/bin/env python
"""Time full-frame rendering (640x480, 256+256 samples) of a random-init network: ms per view, optionally per kernel.
Used to sweep the bring-up knobs of the forward heads kernels in inference (UCSA_FWD_COLOR_CTAS / UCSA_FWD_SEM_CTAS /
UCSA_FWD_OVERLAP), which save no activations there.

    UCSA_FWD_COLOR_CTAS=3 UCSA_FWD_SEM_CTAS=3 python scripts/render_probe.py [--kernels]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ucsa_neural_rendering_b200 import _lib, ops  # noqa: E402
from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork  # noqa: E402
from ucsa_neural_rendering_b200.scene import SyntheticScene  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    scene = SyntheticScene(seed=0, device=dev)
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                              num_semantic_classes=40).to(dev).eval()
    if os.environ.get("STAGE_CHUNK"):
        net.stage_chunk = int(os.environ["STAGE_CHUNK"])
    pix = torch.arange(scene.W * scene.H, device=dev)
    with torch.no_grad():
        vo, vd, vdn = scene.rays(0, pix)
        args = dict(direction_norms=vdn.view(1, -1, 1), staged=True, bg_color=None, perturb=False, seed=99)
        for _ in range(2):
            net.render(vo[None], vd[None], **args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        n = 4
        for _ in range(n):
            out = net.render(vo[None], vd[None], **args)
            ops.label_epilogue(out["semantics"][0], out["image"][0])
        e1.record()
        torch.cuda.synchronize()
        res = {"ms_per_view": e0.elapsed_time(e1) / n, "stage_chunk": net.stage_chunk,
               "peak_alloc_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2),
               "knobs": {k: v for k, v in os.environ.items() if k.startswith("UCSA_")}}
        if "--kernels" in sys.argv:
            names = {"ucsa_density_fwd", "ucsa_heads_fwd", "ucsa_resample_merge", "ucsa_weights_compact",
                     "ucsa_sample_coarse", "ucsa_near_far_from_aabb", "ucsa_label_epilogue"}
            _lib.stats.reset()
            _lib.stats.timed = set(names)
            net.render(vo[None], vd[None], **args)
            torch.cuda.synchronize()
            res["kernel_ms_per_view"] = {k: round(sum(_lib.stats.elapsed_ms(k)), 3) for k in names}
            _lib.stats.timed = set()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
