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
