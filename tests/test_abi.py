"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, and exports every symbol
include/ucsa_nerf.h declares; host-only entry points behave.  No kernel is launched here."""
import ctypes

import numpy as np
import pytest

from oracle import tcnn_spec as spec
from ucsa_neural_rendering_b200 import _lib, build


@pytest.fixture(scope="module")
def handle():
    build.build_library()
    return _lib.lib()


def test_header_symbols_exported(handle):
    decls = _lib.parse_header()
    assert len(decls) >= 26
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in decls:
        assert hasattr(raw, name), f"{name} declared in ucsa_nerf.h but not exported"


def test_abi_version_and_error_string(handle):
    assert handle.ucsa_abi_version() == 1
    rc = handle.ucsa_grid_desc_init(0.5, 16, 19, ctypes.byref(_lib.GridDesc()))
    assert rc == -1
    assert b"unsupported" in handle.ucsa_last_error_string()


def test_null_pointer_is_rejected_not_dereferenced(handle):
    assert handle.ucsa_near_far_from_aabb(None, None, None, 4, 0.2, None, None, None) == -1
    assert b"null" in handle.ucsa_last_error_string()


def test_grid_geometry_matches_oracle(handle):
    from ucsa_neural_rendering_b200 import ops

    for bound in (1, 2, 4, 8):
        g = ops.make_grid_desc(bound)
        t = spec.level_table(bound)
        assert np.array_equal(np.array(list(g.scale), dtype=np.float32), t["scale"])
        for field in ("res", "entries", "offset"):
            assert list(getattr(g, field)) == [int(v) for v in t[field]]
        assert list(g.hashed) == [int(v) for v in t["hashed"]]
        assert g.total_entries == t["total"]


def test_facade_constructs_without_gpu_and_refuses_cpu_render():
    import torch

    from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork

    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1, num_semantic_classes=40)
    names = dict(net.named_parameters())
    assert names["encoder.params"].numel() == 13074912
    assert names["sigma_net.params"].numel() == 3072
    assert names["color_net.params"].numel() == 7168
    assert names["semantics_net.params"].numel() == 4096
    assert set(net.state_dict()) >= {"aabb_train", "aabb_infer", "encoder.params"}
    with pytest.raises(_lib.UcsaError):
        net.render(torch.zeros(1, 4, 3), torch.ones(1, 4, 3), torch.ones(1, 4, 1))


def test_host_side_argument_checks_run_before_any_launch(handle):
    """Entry points validate on the host and return -1 with a message (no GPU needed: they fail before launching)."""
    from ucsa_neural_rendering_b200 import ops

    g = ops.make_grid_desc(4)
    one = ctypes.c_void_p(16)  # a non-null, 16-byte aligned dummy address; never dereferenced on these paths
    # a hashed level whose size is not a power of two (kernels reduce the hash with a mask)
    bad = _lib.GridDesc.from_buffer_copy(g)
    bad.entries[15] = 524288 - 8
    rc = handle.ucsa_hashgrid_indices(one, 4, ctypes.byref(bad), one, None)
    assert rc == -1 and b"grid descriptor" in handle.ucsa_last_error_string()
    # tile-layout activations need T, k0 and k1-k0 to be multiples of 128 in ray mode
    rc = handle.ucsa_density_fwd(None, one, one, one, one, 8, 96, 0, 48, 4.0, one, ctypes.byref(g), one, one, one,
                                 one, one, 1, None)
    assert rc == -1 and b"multiples of 128" in handle.ucsa_last_error_string()
    # the fused exchange: slice bounds in whole float4 groups, rank inside the world
    arr = (ctypes.c_uint64 * 2)(16, 16)
    rc = handle.ucsa_adam_exchange(arr, arr, arr, None, None, None, 2, 0, 2, 8, 0, one, one, 1e-2, 0.9, 0.99, 1e-15,
                                   0.0, 1, None, None, None, 1, None)
    assert rc == -1 and b"multiples of 4" in handle.ucsa_last_error_string()
    rc = handle.ucsa_adam_exchange(arr, arr, arr, None, None, None, 2, 2, 0, 8, 0, one, one, 1e-2, 0.9, 0.99, 1e-15,
                                   0.0, 1, None, None, None, 1, None)
    assert rc == -1 and b"rank < world" in handle.ucsa_last_error_string()
    # the optimizer's overflow check and the loss kernel refuse missing scratch blocks instead of sharing globals
    rc = handle.ucsa_grad_check(one, 16, one, None, None, None)
    assert rc == -1 and b"null" in handle.ucsa_last_error_string()
    rc = handle.ucsa_nerf_loss(one, one, one, one, None, one, one, 4, 40, 1.0, 0.04, 0.1, 1.0, one, one, one, one, None,
                               None)
    assert rc == -1 and b"null" in handle.ucsa_last_error_string()
    # ragged training composite: probabilities or logits (not both), a logit row at least as wide as the class count,
    # [M,2] step sizes read as 8-byte pairs
    args = lambda sem, lg, ld, deltas: (one, one, sem, lg, ld, deltas, one, 64, 4, 40, one, one, one, one, None)  # noqa: E731
    rc = handle.ucsa_composite_rays_train_forward(*args(one, one, 48, one))
    assert rc == -1 and b"not both" in handle.ucsa_last_error_string()
    rc = handle.ucsa_composite_rays_train_forward(*args(None, one, 32, one))
    assert rc == -1 and b"row stride" in handle.ucsa_last_error_string()
    rc = handle.ucsa_composite_rays_train_forward(*args(None, one, 48, ctypes.c_void_p(20)))
    assert rc == -1 and b"8-byte aligned" in handle.ucsa_last_error_string()


def test_tile_layout_helpers_and_launch_accounting():
    from ucsa_neural_rendering_b200 import ops

    assert [ops.tile_rows(n) for n in (0, 1, 128, 129, 4096 * 512)] == [0, 128, 128, 256, 4096 * 512]
    assert ops.density_tiled(256, 256) and ops.density_tiled(128, 0) and not ops.density_tiled(64, 64)
    # ucsa_heads_fwd / ucsa_heads_bwd enqueue two kernels each (colour, semantics); bench.py's gpu_launches uses this
    assert _lib.KERNELS_PER_CALL == {"ucsa_heads_fwd": 2, "ucsa_heads_bwd": 2}
