"""Host-side logic that needs no GPU: workspace caching of the autograd path, the optimizer's argument checks, the
label palette, owner slices of the flat parameter space."""
import gc

import numpy as np
import pytest
import torch

from ucsa_neural_rendering_b200 import pipeline
from ucsa_neural_rendering_b200.labels import default_palette
from ucsa_neural_rendering_b200.optim import FusedAdam


class _Net(torch.nn.Module):
    pass


def test_cached_workspace_is_reused_only_when_its_backward_is_done():
    net = _Net()
    ws1, tok1 = pipeline.cached_workspace(net, 8, 128, 128, 40, torch.device("cpu"), True)
    assert tok1 is not None and ws1.enc is not None and ws1.tiled
    out1 = ws1.image
    # the first call's autograd node (token) is still alive: a second call must get its own workspace
    ws2, tok2 = pipeline.cached_workspace(net, 8, 128, 128, 40, torch.device("cpu"), True)
    assert ws2 is not ws1 and ws2.h.data_ptr() != ws1.h.data_ptr()
    del tok1
    gc.collect()
    # released: the cached one is handed out again, with FRESH output tensors (callers keep the old ones)
    ws3, tok3 = pipeline.cached_workspace(net, 8, 128, 128, 40, torch.device("cpu"), True)
    assert ws3 is ws1 and ws3.image.data_ptr() != out1.data_ptr()
    # inference workspaces (no token) are always reusable and carry no saved-activation buffers
    wi, ti = pipeline.cached_workspace(net, 8, 128, 128, 40, torch.device("cpu"), False)
    wi2, _ = pipeline.cached_workspace(net, 8, 128, 128, 40, torch.device("cpu"), False)
    assert ti is None and wi is wi2 and wi.enc is None
    # other shapes get other workspaces; the cache is bounded per network
    for n in range(9, 9 + 2 * pipeline._WS_CACHE_MAX):
        pipeline.cached_workspace(net, n, 16, 16, 40, torch.device("cpu"), False)
    assert len(pipeline._WS_CACHE[net]) <= pipeline._WS_CACHE_MAX
    del tok2, tok3


def test_fused_adam_checks_its_arguments_and_refuses_cpu_parameters():
    p = torch.nn.Parameter(torch.zeros(8))
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([p], betas=(1.0, 0.99))
    opt = FusedAdam([{"params": [p]}, {"params": [], "weight_decay": 1e-6}], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    assert FusedAdam._step_supports_amp_scaling and not hasattr(opt, "grad_scale")  # GradScaler's protocol
    assert opt.step() is None  # no gradients yet: nothing to do
    p.grad = torch.ones(8)
    with pytest.raises(Exception):
        opt.step()  # CPU parameter: the optimizer has no CPU path


def test_default_palette():
    pal = default_palette(41)
    assert pal.shape == (41, 3) and pal.dtype == np.uint8 and (pal[0] == 0).all()
    assert len({tuple(c) for c in pal.tolist()}) == 41  # all labels distinguishable


def test_occupancy_path_knobs_and_refusals_without_a_gpu():
    """cuda_ray=True facade on the host: the reference's buffers exist, the fused training render steps aside when it is
    switched off or there is nothing to render (the caller then composes the module-level pieces like the reference's
    run_cuda), and no packed-point operator accepts host tensors (there is no CPU fallback)."""
    from ucsa_neural_rendering_b200 import _lib
    from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork

    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=True, density_scale=1, num_semantic_classes=40)
    assert tuple(net.density_grid.shape) == (3, 128, 128, 128) and tuple(net.step_counter.shape) == (16, 2)
    lo, hi = net.wavefront_steps
    assert 1 <= lo <= hi
    pts, rays = torch.zeros(5, 3), torch.zeros(1, 3, dtype=torch.int32)
    assert net.render_packed_train(torch.zeros(0, 3), torch.zeros(0, 3), torch.zeros(0, 2), rays) is None
    net.fused_packed = False
    assert net.render_packed_train(pts, pts, torch.zeros(5, 2), rays) is None
    net.fused_packed = True
    for call in (lambda: net.render_packed_train(pts, pts, torch.zeros(5, 2), rays),
                 lambda: net.forward_packed_train(pts, pts), lambda: net.forward_packed(pts, pts)):
        with pytest.raises(_lib.UcsaError):
            call()


def test_incremental_build_tracks_the_include_closure():
    """An object is rebuilt when its source or anything it includes (transitively, with quotes) changes -- a stale
    object behind an edited header would ship kernels that disagree about a layout."""
    import os

    from ucsa_neural_rendering_b200 import build

    def closure(src):
        deps = set()
        build._closure(os.path.join(build.CSRC, src), deps)
        return {os.path.basename(p) for p in deps}

    header = "ucsa_nerf.h"
    for src in build._sources():
        assert {src, "common.cuh", header} <= closure(src), src
    assert {"grid.cuh", "mlp_umma.cuh"} <= closure("density_tc.cu")
    assert "mlp_umma.cuh" in closure("heads_tc.cu") and "grid.cuh" not in closure("march.cu")
    assert "heads_tc.cu" not in closure("march.cu")
    stamps = {src: build._source_stamp(src) for src in build._sources()}
    assert len(set(stamps.values())) == len(stamps)
