"""Edge cases of the drop-in path on the GPU: odd class counts, rays without any masked-in sample (K = 0), one ray,
ragged tiles, two forward passes before the backward passes (cached workspaces), an overflowing step in the engine."""
import numpy as np
import pytest
import torch

from oracle import live_path

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _net(heads, classes):
    from ucsa_neural_rendering_b200 import build
    from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork

    build.build_library()
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1, num_semantic_classes=classes)
    with torch.no_grad():
        net.encoder.params.copy_(heads.encoder)
        net.sigma_net.params.copy_(heads.sigma_net)
        net.color_net.params.copy_(heads.color_net)
        net.semantics_net.params[:heads.semantics_net.numel()].copy_(heads.semantics_net)
        net.semantics_net.params[heads.semantics_net.numel():].zero_()
    return net.to(DEV)


def _rays(n, seed):
    g = torch.Generator().manual_seed(seed)
    o = (torch.rand(1, n, 3, generator=g) - 0.5) * 2
    d = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1)
    dn = 1 + 0.2 * torch.rand(1, n, 1, generator=g)
    return g, o, d, dn


def _close(name, got, ref, rtol, atol_rel):
    ref = np.asarray(ref)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol_rel * max(np.abs(ref).max(), 1e-12), err_msg=name)


@pytest.mark.parametrize("classes,n", [(1, 37), (33, 129), (48, 1)])
def test_class_counts_and_ragged_ray_counts_against_oracle(classes, n):
    """the semantic output layer is padded to 48 columns inside the kernels: 1, 33 and 48 classes, with 37 / 129 / 1
    rays (partial tiles, a single ray), forward and backward against the oracle"""
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=classes, seed=50 + classes, hash_amp=0.4)
    net = _net(heads, classes).train()
    steps = up = 24  # not multiples of 128: the row-major saved-activation path
    g, o, d, dn = _rays(n, classes)
    t_rand, u = torch.rand(n, steps, generator=g), torch.rand(n, up, generator=g)
    ref = live_path.run(heads, o, d, dn, num_steps=steps, upsample_steps=up, perturb=True, t_rand=t_rand, u=u)
    out = net.render(o.to(DEV), d.to(DEV), direction_norms=dn.to(DEV), perturb=True, num_steps=steps, upsample_steps=up,
                     t_rand=t_rand.to(DEV), u=u.to(DEV))
    assert out["semantics"].shape == (1, n, classes)
    for k in ("depth", "image", "semantics"):
        _close(k, out[k].detach().cpu().numpy(), ref[k].detach().numpy(), rtol=4e-3, atol_rel=2e-3)
    gs = torch.randn(1, n, classes, generator=g)
    ((ref["image"] ** 2).sum() + (ref["semantics"] * gs).sum() + ref["depth"].sum()).backward()
    ((out["image"] ** 2).sum() + (out["semantics"] * gs.to(DEV)).sum() + out["depth"].sum()).backward()
    n_sem = heads.semantics_net.numel()
    for name, p_ref, got in (("sigma", heads.sigma_net, net.sigma_net.params.grad),
                             ("color", heads.color_net, net.color_net.params.grad),
                             ("sem", heads.semantics_net, net.semantics_net.params.grad[:n_sem]),
                             ("hash", heads.encoder, net.encoder.params.grad)):
        _close("grad_" + name, got.cpu().numpy(), p_ref.grad.numpy(), rtol=3e-2, atol_rel=1e-2)
    # rows of the padded output layer that belong to no class receive no gradient
    assert float(net.semantics_net.params.grad[n_sem:].abs().sum()) == 0.0 or classes % 16 == 0


def test_no_sample_passes_the_mask():
    """K = 0: with density_scale = 0 every alpha is 0, no weight passes w > 1e-4, the heads are skipped for every
    sample (network_tcnn_semantics.py:157-158): all three outputs are exactly zero, and the backward pass runs on the
    empty compact list and returns zero gradients."""
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=3, hash_amp=0.3)
    net = _net(heads, 40).train()
    net.density_scale = 0.0
    n = 70
    g, o, d, dn = _rays(n, 9)
    out = net.render(o.to(DEV), d.to(DEV), direction_norms=dn.to(DEV), perturb=True, num_steps=128, upsample_steps=128)
    for k in ("depth", "image", "semantics"):
        assert float(out[k].abs().sum()) == 0.0, k
    (out["image"].sum() + out["semantics"].sum() + out["depth"].sum()).backward()
    for p in net.parameters():
        assert p.grad is not None and float(p.grad.abs().sum()) == 0.0


def test_two_forwards_before_the_backwards_use_separate_workspaces():
    """render() caches its workspace per shape; a second forward of the same shape before the first backward must not
    overwrite what the first backward needs"""
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=8, hash_amp=0.4)
    net = _net(heads, 40).train()
    n = 64
    g, o, d, dn = _rays(n, 4)
    args = dict(direction_norms=dn.to(DEV), perturb=True, num_steps=32, upsample_steps=32)
    o1, d1 = o.to(DEV), d.to(DEV)
    o2 = (o * 0.5).to(DEV)

    def grads_of(loss):
        net.zero_grad(set_to_none=True)
        loss.backward()
        return [p.grad.clone() for p in net.parameters()]

    a = net.render(o1, d1, seed=11, **args)
    ga = grads_of(a["image"].sum() + a["semantics"].sum())
    b = net.render(o2, d1, seed=12, **args)
    gb = grads_of(b["image"].sum() + b["semantics"].sum())
    # interleaved: both graphs alive at once
    a2 = net.render(o1, d1, seed=11, **args)
    b2 = net.render(o2, d1, seed=12, **args)
    torch.testing.assert_close(a2["image"], a["image"], rtol=1e-5, atol=1e-6)
    assert a2["image"].data_ptr() != b2["image"].data_ptr()
    gb2 = grads_of(b2["image"].sum() + b2["semantics"].sum())
    ga2 = grads_of(a2["image"].sum() + a2["semantics"].sum())
    for x, y in zip(ga + gb, ga2 + gb2):  # equal up to the order of the atomics
        torch.testing.assert_close(x, y, rtol=1e-3, atol=1e-4 * float(y.abs().max()) + 1e-12)
    with pytest.raises(RuntimeError):
        (a2["image"].sum()).backward()  # the node was differentiated and released its workspace


def test_engine_skips_a_step_whose_gradient_overflows():
    """ADVICE (round 1): a single fp16 overflow must not poison moments / masters / fp16 copies.  A poisoned target
    makes the loss and hence the flat gradient non-finite: the engine skips that step, parameters and moments are
    untouched, Adam's step count does not advance, and training continues."""
    from ucsa_neural_rendering_b200.engine import TrainEngine

    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=2, hash_amp=0.3)
    net = _net(heads, 40).train()
    n = 256
    eng = TrainEngine(net, n, num_steps=32, upsample_steps=32, one_m_to_scene_uom=0.6, seed=1, use_graph=True)
    g, o, d, dn = _rays(n, 6)
    batch = [o[0].to(DEV), d[0].to(DEV), dn[0, :, 0].to(DEV), torch.rand(n, 3, generator=g).half().to(DEV),
             torch.randint(0, 40, (n,), generator=g).to(DEV), (torch.rand(n, generator=g) * 3).to(DEV)]
    for _ in range(2):
        eng.train_step(*batch)
    torch.cuda.synchronize()
    before = [m.params.detach().clone() for m, _ in eng.groups] + [t.clone() for t in eng.exp_avg + eng.exp_avg_sq]
    halves = [m.half_params().clone() for m, _ in eng.groups]
    bad = [t.clone() for t in batch]
    bad[3][0, 0] = float("inf")
    loss = eng.train_step(*bad)
    torch.cuda.synchronize()
    assert not torch.isfinite(loss[0]) and eng.skipped_steps == 1 and float(eng.found_inf) == 1.0
    after = [m.params.detach() for m, _ in eng.groups] + eng.exp_avg + eng.exp_avg_sq
    for x, y in zip(before, after):
        assert torch.equal(x, y), "a skipped step must leave masters and moments untouched"
    for x, (m, _) in zip(halves, eng.groups):
        assert torch.equal(x, m.half_params())
    loss = eng.train_step(*batch)
    torch.cuda.synchronize()
    assert torch.isfinite(loss[0]) and float(eng.found_inf) == 0.0
    assert int(eng.step_dev) == 4 and eng.skipped_steps == 1  # Adam has taken 3 steps
    assert not torch.equal(before[0], eng.groups[0][0].params.detach())


def test_opt_in_kernel_variants_pass_the_heads_parity_test():
    """The bring-up variants that stay in the library behind environment knobs (warp-specialised two-slot colour
    forward kernel, concurrent backward heads kernels, sequential forward heads kernels) must keep passing the heads
    parity test; the knobs are read once per process, hence a subprocess per variant."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env in ({"UCSA_HEADS_WS": "1"}, {"UCSA_BWD_OVERLAP": "1"}, {"UCSA_FWD_OVERLAP": "0"}):
        res = subprocess.run(
            [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider",
             os.path.join(root, "tests", "test_gpu_render.py"), "-k",
             "fused_heads_match_cuda_core or train_small_golden or fused_compositing_isolated"],
            env=dict(os.environ, **env), capture_output=True, text=True, timeout=600, cwd=root)
        assert res.returncode == 0, (env, res.stdout[-1500:], res.stderr[-500:])
