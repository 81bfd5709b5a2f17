"""End-to-end parity of the drop-in SemanticNeRFNetwork.render() (fused CUDA pipeline, called through the
reference's own API) against the golden fixtures frozen from the reference modules and against the oracle.
GPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import live_path

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _net_from_oracle(heads, classes=40):
    from ucsa_neural_rendering_b200 import build
    from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork

    build.build_library()
    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                              num_semantic_classes=classes)
    with torch.no_grad():
        net.encoder.params.copy_(heads.encoder)
        net.sigma_net.params.copy_(heads.sigma_net)
        net.color_net.params.copy_(heads.color_net)
        net.semantics_net.params.copy_(heads.semantics_net)
    return net.to(DEV)


def _oracle_heads(g):
    return live_path.OracleHeads(bound=4, num_semantic_classes=int(g["cfg"][7]), seed=int(g["cfg"][8]),
                                 hash_amp=float(g["hash_amp"]))


def _close(name, got, ref, rtol, atol_rel):
    ref = np.asarray(ref)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol_rel * max(np.abs(ref).max(), 1e-12), err_msg=name)


@pytest.mark.parametrize("generic", [False, True])
def test_train_small_golden(golden_dir, generic):
    """Forward + backward of one training-mode render against the reference's own outputs / gradients."""
    g = _load(golden_dir, "train_small")
    n, steps, up = [int(v) for v in g["cfg"][:3]]
    net = _net_from_oracle(_oracle_heads(g))
    net.train()
    kw = dict(generic=True) if generic else {}
    out = net.render(_t(g["rays_o"]).to(DEV), _t(g["rays_d"]).to(DEV), direction_norms=_t(g["direction_norms"]).to(DEV),
                     staged=False, bg_color=None, perturb=True, num_steps=steps, upsample_steps=up,
                     t_rand=_t(g["t_rand"]).to(DEV), u=_t(g["u"]).to(DEV), **kw)
    assert out["image"].dtype == torch.float32 and out["image"].shape == (1, n, 3)
    assert out["depth"].shape == (1, n) and out["semantics"].shape == (1, n, 40)
    # fp16 encoding / MLP chain => 2e-3
    for k in ("depth", "image", "semantics"):
        _close(k, out[k].detach().cpu().numpy(), g[k], rtol=4e-3, atol_rel=2e-3)
    loss = (out["image"] * _t(g["g_image"]).to(DEV)).sum() + (out["depth"] * _t(g["g_depth"]).to(DEV)).sum() \
        + (out["semantics"] * _t(g["g_semantics"]).to(DEV)).sum()
    loss.backward()
    for name, mod in (("sigma_net", net.sigma_net), ("color_net", net.color_net),
                      ("semantics_net", net.semantics_net)):
        _close("grad_" + name, mod.params.grad.cpu().numpy(), g["grad_" + name], rtol=3e-2, atol_rel=1e-2)
    gh = net.encoder.params.grad.cpu()
    _close("grad_hash", gh[_t(g["grad_hash_idx"])].numpy(), g["grad_hash_val"], rtol=3e-2, atol_rel=1e-2)
    assert abs(float(gh.double().abs().sum()) - float(g["grad_hash_abs"])) < 2e-2 * float(g["grad_hash_abs"])


def test_infer_staged_golden(golden_dir):
    """staged=True chunk loop (ragged tail, rays that miss the box) in eval mode."""
    g = _load(golden_dir, "infer_staged")
    n, steps, up = [int(v) for v in g["cfg"][:3]]
    net = _net_from_oracle(_oracle_heads(g))
    net.eval()
    args = dict(direction_norms=_t(g["direction_norms"]).to(DEV), staged=True, max_ray_batch=int(g["cfg"][6]),
                bg_color=1, perturb=False, num_steps=steps, upsample_steps=up, u=_t(g["u"]).to(DEV))
    net.stage_chunk = None  # the reference's chunk loop exactly: max_ray_batch rays per pass, ragged tail
    with torch.no_grad():
        out = net.render(_t(g["rays_o"]).to(DEV), _t(g["rays_d"]).to(DEV), **args)
    for k in ("depth", "image", "semantics"):
        _close(k, out[k].cpu().numpy(), g[k], rtol=4e-3, atol_rel=2e-3)
    # default: without gradients the network re-chunks to >= 65536 rays; the result does not depend on the chunking
    net.stage_chunk = 65536
    with torch.no_grad():
        out2 = net.render(_t(g["rays_o"]).to(DEV), _t(g["rays_d"]).to(DEV), **args)
    for k in out:
        torch.testing.assert_close(out2[k], out[k], rtol=1e-5, atol=1e-6)
    # inference must not allocate (or write) the activations a backward pass would need: torch reports
    # needs_input_grad = True under no_grad as well, the grad mode has to reach the fused node explicitly
    from ucsa_neural_rendering_b200 import pipeline

    cached = pipeline._WS_CACHE[net]
    assert cached and all(not ws.need_grad and ws.enc is None and ws.hc1 is None for ws in cached.values())


def test_fused_equals_generic_path_at_reference_sizes():
    """256 + 256 samples (the reference defaults): the fused pipeline and the method-by-method path agree."""
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=5, hash_amp=0.3)
    net = _net_from_oracle(heads)
    net.train()
    n = 512
    g = torch.Generator().manual_seed(17)
    o = ((torch.rand(1, n, 3, generator=g) - 0.5) * 2).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1).to(DEV)
    dn = (1 + 0.2 * torch.rand(1, n, 1, generator=g)).to(DEV)
    t_rand = torch.rand(n, 256, generator=g).to(DEV)
    u = torch.rand(n, 256, generator=g).to(DEV)
    outs = []
    grads = []
    for generic in (False, True):
        net.zero_grad(set_to_none=True)
        kw = dict(generic=True) if generic else {}
        out = net.render(o, d, direction_norms=dn, staged=False, perturb=True, t_rand=t_rand, u=u, **kw)
        (out["image"].sum() + out["depth"].sum() + (out["semantics"] ** 2).sum()).backward()
        outs.append({k: v.detach() for k, v in out.items()})
        grads.append([p.grad.clone() for p in net.parameters()])
    for k in outs[0]:
        torch.testing.assert_close(outs[0][k], outs[1][k], rtol=2e-3, atol=2e-3)
    for a, b in zip(*grads):
        scale = float(b.abs().max())
        torch.testing.assert_close(a, b, rtol=3e-2, atol=1e-2 * scale)


def test_oracle_on_fresh_inputs():
    """A second, larger seeded case straight against the CPU oracle (no fixture)."""
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=77, hash_amp=0.4)
    net = _net_from_oracle(heads)
    net.train()
    n, steps, up = 200, 64, 64
    g = torch.Generator().manual_seed(4)
    o = (torch.rand(1, n, 3, generator=g) - 0.5) * 3
    d = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1)
    dn = 1 + 0.3 * torch.rand(1, n, 1, generator=g)
    t_rand = torch.rand(n, steps, generator=g)
    u = torch.rand(n, up, generator=g)
    ref = live_path.run(heads, o, d, dn, num_steps=steps, upsample_steps=up, perturb=True, t_rand=t_rand, u=u)
    out = net.render(o.to(DEV), d.to(DEV), direction_norms=dn.to(DEV), perturb=True, num_steps=steps,
                     upsample_steps=up, t_rand=t_rand.to(DEV), u=u.to(DEV))
    for k in ("depth", "image", "semantics"):
        _close(k, out[k].detach().cpu().numpy(), ref[k].detach().numpy(), rtol=4e-3, atol_rel=2e-3)
    gi = torch.randn(1, n, 3, generator=g)
    gs = torch.randn(1, n, 40, generator=g)
    ((ref["image"] * gi).sum() + (ref["semantics"] * gs).sum() + ref["depth"].sum()).backward()
    ((out["image"] * gi.to(DEV)).sum() + (out["semantics"] * gs.to(DEV)).sum() + out["depth"].sum()).backward()
    for name, p_ref, mod in (("sigma", heads.sigma_net, net.sigma_net), ("color", heads.color_net, net.color_net),
                             ("sem", heads.semantics_net, net.semantics_net), ("hash", heads.encoder, net.encoder)):
        _close("grad_" + name, mod.params.grad.cpu().numpy(), p_ref.grad.numpy(), rtol=3e-2, atol_rel=1e-2)


def test_module_level_api_matches_reference_contract():
    """density()/color()/semantics()/forward(): shapes, dtypes and masked semantics of the reference network."""
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=3, hash_amp=0.4)
    net = _net_from_oracle(heads)
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(777, 3, generator=g) - 0.5) * 8
    d = torch.nn.functional.normalize(torch.randn(777, 3, generator=g), dim=-1)
    mask = torch.rand(777, generator=g) > 0.5
    dens = net.density(x.to(DEV))
    ref = heads.density(x)
    assert dens["sigma"].dtype == torch.float32 and dens["geo_feat"].dtype == torch.float16
    _close("sigma", dens["sigma"].detach().cpu().numpy(), ref["sigma"].detach().numpy(), 4e-3, 2e-3)
    _close("geo", dens["geo_feat"].float().detach().cpu().numpy(), ref["geo_feat"].detach().numpy(), 4e-3, 2e-3)
    rgb = net.color(x.to(DEV), d.to(DEV), mask=mask.to(DEV), geo_feat=dens["geo_feat"])
    rgb_ref = heads.color(x, d, mask=mask, geo_feat=ref["geo_feat"])
    assert rgb.shape == (777, 3) and rgb.dtype == torch.float32 and (rgb[~mask.to(DEV)] == 0).all()
    _close("rgb", rgb.detach().cpu().numpy(), rgb_ref.detach().numpy(), 4e-3, 2e-3)
    sem = net.semantics(x.to(DEV), d.to(DEV), mask=mask.to(DEV), geo_feat=dens["geo_feat"])
    sem_ref = heads.semantics(x, d, mask=mask, geo_feat=ref["geo_feat"])
    assert sem.shape == (777, 40) and (sem[~mask.to(DEV)] == 0).all()
    _close("sem", sem.detach().cpu().numpy(), sem_ref.detach().numpy(), 4e-3, 2e-3)
    empty = torch.zeros(777, dtype=torch.bool, device=DEV)
    assert (net.color(x.to(DEV), d.to(DEV), mask=empty, geo_feat=dens["geo_feat"]) == 0).all()
    s2, c2, p2 = net(x.to(DEV), d.to(DEV))
    s_ref, c_ref, p_ref = heads(x, d)
    _close("fwd_color", c2.float().detach().cpu().numpy(), c_ref.detach().numpy(), 4e-3, 2e-3)
    _close("fwd_sem", p2.detach().cpu().numpy(), p_ref.detach().numpy(), 4e-3, 2e-3)
    # autograd through the module-level path reaches every parameter group
    (s2.sum() + c2.float().sum() + (p2 ** 2).sum()).backward()
    for p in net.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0


def test_optimizer_step_changes_render_and_state_dict_round_trips():
    """Two Adam groups as at joint_train_lightning_net.py:897-919; fp16 working copies follow the masters."""
    from ucsa_neural_rendering_b200.nerf import SemanticNeRFNetwork

    net = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                              num_semantic_classes=40).to(DEV)
    opt = torch.optim.Adam([
        {"name": "encoding", "params": list(net.encoder.parameters())},
        {"name": "net", "params": list(net.sigma_net.parameters()) + list(net.color_net.parameters())
         + list(net.semantics_net.parameters()), "weight_decay": 1e-6},
    ], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    n = 256
    g = torch.Generator().manual_seed(1)
    o = ((torch.rand(1, n, 3, generator=g) - 0.5)).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1).to(DEV)
    dn = torch.ones(1, n, 1, device=DEV)
    target = torch.rand(1, n, 3, generator=g).to(DEV)
    scaler = torch.amp.GradScaler("cuda", enabled=True)
    losses = []
    net.train()
    for _ in range(12):
        opt.zero_grad()
        with torch.autocast("cuda", enabled=True):
            out = net.render(o, d, direction_norms=dn, perturb=True, num_steps=32, upsample_steps=32, seed=7)
            loss = ((out["image"] - target) ** 2).mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
    sd = net.state_dict()
    net2 = SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                               num_semantic_classes=40).to(DEV)
    net2.load_state_dict(sd)
    net.eval(), net2.eval()
    with torch.no_grad():
        a = net.render(o, d, direction_norms=dn, staged=True, perturb=False, seed=9)
        b = net2.render(o, d, direction_norms=dn, staged=True, perturb=False, seed=9)
    for k in a:  # the fused compositing adds per-tile partial sums with atomics: equal up to fp32 summation order
        torch.testing.assert_close(a[k], b[k], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,t", [(300, 48), (6000, 128)], ids=["one-tile-per-cta", "many-tiles-per-cta"])
def test_fused_heads_match_cuda_core_heads_plus_composite_kernels(n, t):
    """tcgen05 heads with fused compositing (production) vs CUDA-core heads + stand-alone composite kernels.
    The large case gives every persistent CTA several 128-row tiles (barrier phases, tile re-use, TMEM accumulation
    of the weight gradients across tiles)."""
    from ucsa_neural_rendering_b200 import ops

    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=9, hash_amp=0.4)
    net = _net_from_oracle(heads)
    c = 40
    g = torch.Generator().manual_seed(8)
    f32 = dict(dtype=torch.float32, device=DEV)
    f16 = dict(dtype=torch.float16, device=DEV)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(DEV)
    h = (torch.randn(n, t, 16, generator=g) * 0.5).half().to(DEV)
    w_all = torch.rand(n, t, generator=g) ** 6
    w_all[5] = 0  # a ray without any masked-in sample
    z = torch.sort(torch.rand(n, t, generator=g) * 4 + 0.2, dim=1).values.to(DEV)
    w_all = w_all.to(DEV)
    dn = (1 + 0.2 * torch.rand(n, generator=g)).to(DEV)
    cnt = (w_all > 1e-4).sum(1).int()
    off = torch.zeros(n + 1, dtype=torch.int32, device=DEV)
    ops.scan_counts(cnt, off)
    k = int(off[-1])
    k_max = n * t
    sel = torch.empty(k_max, dtype=torch.int32, device=DEV)
    w_sel, z_sel = torch.empty(k_max, **f32), torch.empty(k_max, **f32)
    ops.compact_masked(w_all, z, None, off, sel, w_sel, z_sel)
    w_col, w_sem = net.color_net.half_params(), net.semantics_net.half_params()

    rgb1, log1 = torch.zeros(k_max, 3, **f32), torch.zeros(k_max, 48, **f16)
    a1, a2, a3 = (torch.zeros(ops.tile_rows(k_max), 64, **f16) for _ in range(3))  # tile-layout buffers
    img1, sem1 = torch.zeros(n, 3, **f32), torch.zeros(n, c, **f32)
    ops.heads_fwd(sel, off, n, t, k_max, d, h, w_col, w_sem, c, rgb1, log1, a1, a2, a3, w_sel=w_sel, image=img1,
                  semantics=sem1)
    rgb2, log2 = torch.zeros(k_max, 3, **f32), torch.zeros(k_max, 48, **f16)
    b1, b2, b3 = (torch.zeros(k_max, 64, **f16) for _ in range(3))
    ops.heads_fwd_simt(sel, off, n, t, k_max, d, h, w_col, w_sem, c, rgb2, log2, b1, b2, b3)
    img2, sem2 = torch.empty(n, 3, **f32), torch.empty(n, c, **f32)
    ops.composite_fwd(off, w_sel, rgb2, log2, n, c, img2, sem2)
    torch.testing.assert_close(rgb1[:k], rgb2[:k], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(log1[:k].float(), log2[:k].float(), rtol=2e-3, atol=4e-3)
    torch.testing.assert_close(img1, img2, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(sem1, sem2, rtol=2e-3, atol=2e-3)
    assert float(sem1[5].abs().sum()) == 0 and float(img1[5].abs().sum()) == 0

    gi, gd, gs = torch.randn(n, 3, generator=g).to(DEV), torch.randn(n, generator=g).to(DEV), \
        torch.randn(n, c, generator=g).to(DEV)
    scale = 64.0
    dh1, dw1 = torch.zeros(n, t, 16, **f16), torch.zeros(k_max, **f32)
    gc1, gs1 = torch.zeros(ops.COLOR_PARAMS, **f32), torch.zeros(ops.SEM_PARAMS, **f32)
    ops.heads_bwd(sel, off, n, t, k_max, d, h, w_col, w_sem, c, rgb1, a1, a2, a3, w_sel, z_sel, gi, gd, gs, dn,
                  scale, dh1, dw1, gc1, gs1)
    d_rgb, d_log, dw2 = torch.zeros(k_max, 3, **f32), torch.zeros(k_max, 48, **f32), torch.zeros(k_max, **f32)
    ops.composite_bwd(off, sel, w_sel, z_sel, rgb2, log2, gi, gd, gs, dn, n, c, d_rgb, d_log, dw2)
    dh2 = torch.zeros(n, t, 16, **f16)
    gc2, gs2 = torch.zeros(ops.COLOR_PARAMS, **f32), torch.zeros(ops.SEM_PARAMS, **f32)
    ops.heads_bwd_simt(sel, off, n, t, k_max, d, h, w_col, w_sem, c, rgb2, b1, b2, b3, d_rgb, d_log, scale, dh2, gc2,
                       gs2)
    torch.testing.assert_close(dw1[:k], dw2[:k], rtol=2e-3, atol=2e-3 * float(dw2[:k].abs().max()))
    torch.testing.assert_close(gc1, gc2, rtol=2e-2, atol=5e-3 * float(gc2.abs().max()))
    torch.testing.assert_close(gs1, gs2, rtol=2e-2, atol=5e-3 * float(gs2.abs().max()))
    m = (w_all > 1e-4).view(-1)
    a = dh1.view(-1, 16)[m][:, 1:].float()
    b = dh2.view(-1, 16)[m][:, 1:].float()
    # a hidden unit whose fp16 pre-activation is +0 in one implementation and -tiny in the other flips its ReLU mask
    # for that row: tolerate a 1e-5 fraction of such elements, hold everything else to the tolerance
    bad = (a - b).abs() > 2e-2 * b.abs() + 1e-2 * float(b.abs().max())
    assert float(bad.float().mean()) < 1e-5, float(bad.float().mean())
    # concurrent form: the semantic kernel writes its share of dL/dgeo_feat to its own buffer and the two kernels run
    # side by side; half(colour + semantic) must reproduce the in-place sum bit for bit, everything else as before
    dh3, dhs = torch.zeros(n, t, 16, **f16), torch.zeros(n, t, 16, **f16)
    dw3 = torch.zeros(k_max, **f32)
    gc3, gs3 = torch.zeros(ops.COLOR_PARAMS, **f32), torch.zeros(ops.SEM_PARAMS, **f32)
    ops.heads_bwd(sel, off, n, t, k_max, d, h, w_col, w_sem, c, rgb1, a1, a2, a3, w_sel, z_sel, gi, gd, gs, dn,
                  scale, dh3, dw3, gc3, gs3, dh_sem=dhs)
    torch.cuda.synchronize()
    summed = (dh3.float() + dhs.float()).half()
    assert torch.equal(summed.view(-1, 16)[m], dh1.view(-1, 16)[m])
    assert torch.equal(dw3[:k], dw1[:k])
    torch.testing.assert_close(gc3, gc1, rtol=1e-4, atol=1e-5 * float(gc1.abs().max()))
    torch.testing.assert_close(gs3, gs1, rtol=1e-4, atol=1e-5 * float(gs1.abs().max()))


def test_config2_full_size_properties():
    """BASELINE config 2 sizes (4096 rays x 256+256 samples, 40 classes) through size-independent properties:
    determinism for a fixed seed, value ranges, the opacity checksum linking semantics and depth, and linearity of the
    backward pass (the oracle needs minutes at this size)."""
    from ucsa_neural_rendering_b200 import ops
    from ucsa_neural_rendering_b200.scene import SyntheticScene

    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=3, hash_amp=0.3)
    net = _net_from_oracle(heads)
    net.train()
    scene = SyntheticScene(seed=0, device=DEV)
    n = 4096
    g = torch.Generator(device=DEV).manual_seed(2)
    pix = torch.randint(0, scene.W * scene.H, (n,), device=DEV, generator=g)
    o, d, dn = scene.rays(3, pix)
    args = dict(direction_norms=dn.view(1, n, 1), staged=False, perturb=True, seed=77)

    def run(weights):
        net.zero_grad(set_to_none=True)
        out = net.render(o[None], d[None], **args)
        loss = sum(w * out[k].sum() for k, w in weights.items())
        loss.backward()
        return {k: v.detach()[0] for k, v in out.items()}, torch.cat([p.grad.reshape(-1) for p in net.parameters()])

    out1, g_img = run({"image": 1.0})
    out2, g_dep = run({"depth": 1.0})
    out3, g_mix = run({"image": 1.0, "depth": 2.0})
    for k in out1:  # same seed, same samples: equal up to the order of the fused compositing's atomics
        assert torch.isfinite(out1[k]).all()
        torch.testing.assert_close(out1[k], out2[k], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out1[k], out3[k], rtol=1e-5, atol=1e-6)
    image, sem, depth = out1["image"], out1["semantics"], out1["depth"]
    assert float(image.min()) >= 0 and float(image.max()) <= 1 + 1e-4
    assert float(sem.min()) >= 0
    opacity = sem.sum(-1)  # soft-max rows sum to one: sum_c semantics = sum of the masked-in weights
    assert float(opacity.max()) <= 1 + 1e-4 and float(opacity.mean()) > 0.5
    nears, fars = ops.near_far_from_aabb(o, d, net.aabb_train)
    wz = depth * dn.view(-1)  # = sum of masked-in w * z, which lies between near * opacity and far * opacity
    assert bool((wz <= fars * opacity * (1 + 1e-4) + 1e-4).all()) and bool((wz >= nears * opacity * (1 - 1e-4) - 1e-4).all())
    # backward is linear in the output gradients
    assert torch.isfinite(g_mix).all()
    want = g_img + 2.0 * g_dep
    scale = float(want.abs().max())
    bad = (g_mix - want).abs() > 3e-2 * want.abs() + 1e-2 * scale
    assert float(bad.float().mean()) < 1e-5, float(bad.float().mean())
    assert float(g_img.abs().max()) > 0 and float(g_dep.abs().max()) > 0


def _scene_batch(n, seed, view=3):
    from ucsa_neural_rendering_b200.scene import SyntheticScene

    scene = SyntheticScene(seed=0)
    g = torch.Generator().manual_seed(seed)
    pix = torch.randint(0, scene.W * scene.H, (n,), generator=g)
    o, d, dn = scene.rays(view, pix)
    rgb, depth, label = scene.ground_truth(o, d, dn)
    return scene, g, o, d, dn, rgb, depth, label


def test_oracle_at_reference_sample_counts_render_and_engine():
    """256 + 256 samples per ray (renderer_semantics.py:127-128), rays of the benchmark's synthetic scene: the fused
    render() AND the TrainEngine's forward/backward (the path bench.py times) against the CPU oracle, outputs at 2e-3,
    gradients at 3e-2 (fp16 backward chain)."""
    from oracle.losses import nerf_losses as oracle_losses
    from ucsa_neural_rendering_b200.engine import TrainEngine

    n, steps, up = 96, 256, 256
    scene, g, o, d, dn, rgb, depth, label = _scene_batch(n, 21)
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=40, seed=41, hash_amp=0.3)
    net = _net_from_oracle(heads)
    net.train()
    t_rand = torch.rand(n, steps, generator=g)
    u = torch.rand(n, up, generator=g)
    ref = live_path.run(heads, o[None], d[None], dn[None], num_steps=steps, upsample_steps=up, perturb=True,
                        t_rand=t_rand, u=u, return_aux=True)
    ref_aux = ref.pop("aux")
    ref_total, ref_parts = oracle_losses(ref, rgb[None], label[None], depth[None], scene.one_m_to_scene_uom)
    ref_total.backward()
    ref_grads = [heads.encoder.grad, heads.sigma_net.grad, heads.color_net.grad, heads.semantics_net.grad]

    # (1) the drop-in API
    out = net.render(o[None].to(DEV), d[None].to(DEV), direction_norms=dn[None].to(DEV), staged=False, perturb=True,
                     t_rand=t_rand.to(DEV), u=u.to(DEV))
    for k in ("depth", "image", "semantics"):
        _close(k, out[k].detach().cpu().numpy(), ref[k].detach().numpy(), rtol=4e-3, atol_rel=2e-3)

    # (2) the engine: same kernels on a static workspace + fused loss kernel, gradients in one flat buffer
    eng = TrainEngine(net, n, num_steps=steps, upsample_steps=up, one_m_to_scene_uom=scene.one_m_to_scene_uom,
                      use_graph=False)
    eng.t_rand, eng.u = t_rand.to(DEV), u.to(DEV)
    eng.load_batch(o.to(DEV), d.to(DEV), dn.to(DEV), rgb.half().to(DEV), label.to(DEV), depth.to(DEV))
    eng._forward_backward()
    torch.cuda.synchronize()
    for k, got in (("depth", eng.ws.depth), ("image", eng.ws.image), ("semantics", eng.ws.semantics)):
        _close("engine_" + k, got.cpu().numpy(), ref[k].detach().numpy()[0], rtol=4e-3, atol_rel=2e-3)
    loss = eng.loss.cpu().numpy()
    # gt colours are fp16 on the engine side (batch["img_fp16"]): 5e-4 of slack on the colour term
    np.testing.assert_allclose(loss[0], float(ref_total), rtol=4e-3)
    np.testing.assert_allclose(loss[1:], [float(x) for x in ref_parts], rtol=4e-3, atol=1e-5)
    for name, got, want in zip(("hash", "sigma", "color", "sem"), eng.grads, ref_grads):
        _close("engine_grad_" + name, got.cpu().numpy(), want.numpy(), rtol=3e-2, atol_rel=1e-2)
    # Importance samples: inverse-CDF sampling is ill-conditioned inside near-empty bins, where the kernel and torch may
    # place a sample a fraction of a bin apart (tests/test_gpu_kernels.py::test_resample_merge allows 3 % of them).  The
    # rays that contain such samples must still render within tolerance -- checked here on exactly those rays.
    z_fine = eng.ws.z_cat[:, steps:].cpu()
    z_ref = torch.sort(ref_aux["z_new"], dim=1).values
    off = ((z_fine - z_ref).abs() / z_ref.abs().clamp_min(1e-3)) > 2e-5
    assert float(off.float().mean()) < 0.03
    rays_off = off.any(dim=1)
    if bool(rays_off.any()):
        for k, got in (("depth", eng.ws.depth), ("image", eng.ws.image), ("semantics", eng.ws.semantics)):
            a, b = got.cpu().numpy()[rays_off.numpy()], ref[k].detach().numpy()[0][rays_off.numpy()]
            np.testing.assert_allclose(a, b, rtol=4e-3, atol=2e-3 * max(np.abs(ref[k].detach().numpy()).max(), 1e-12),
                                       err_msg="rays with displaced importance samples: " + k)


def test_fused_compositing_isolated_to_1e5():
    """The compositing fused into the heads kernels, isolated from the fp16 MLP error: at config-2 size (4096 rays x
    256+256 samples) the kernel's own per-row colours / logits / weights are re-summed in float64 and compared with the
    kernel's image / semantics / depth at 1e-5 (renderer_semantics.py:269-285); likewise dL/dw of the backward."""
    from ucsa_neural_rendering_b200 import ops, pipeline

    n, tc, tf, c = 4096, 256, 256, 40
    scene, g, o, d, dn, *_ = _scene_batch(n, 5)
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=c, seed=13, hash_amp=0.3)
    net = _net_from_oracle(heads)
    net.train()
    o, d, dn = o.to(DEV), d.to(DEV), dn.view(-1).to(DEV)
    ws = pipeline.RenderWorkspace(n, tc, tf, c, DEV, need_grad=True)
    pipeline.forward_chain(net, ws, o, d, dn, net.aabb_train, perturb=True, seed=3)
    torch.cuda.synchronize()
    off = ws.ray_off.long()
    k = int(off[-1])
    counts = off[1:] - off[:-1]
    assert 0.2 * n * (tc + tf) < k <= n * (tc + tf)
    ray = torch.repeat_interleave(torch.arange(n, device=DEV), counts)
    # the same kernels once more, this time also writing the per-row logits
    logits = torch.zeros(ws.k_max, 48, dtype=torch.float16, device=DEV)
    rgb = torch.zeros(ws.k_max, 3, device=DEV)
    image, sem = torch.zeros(n, 3, device=DEV), torch.zeros(n, c, device=DEV)
    ops.heads_fwd(ws.sel, ws.ray_off, n, tc + tf, ws.k_max, d, ws.h, net.color_net.half_params(),
                  net.semantics_net.half_params(), c, rgb, logits, ws.hc1, ws.hc2, ws.hs, w_sel=ws.w_sel, image=image,
                  semantics=sem)
    torch.cuda.synchronize()
    assert torch.equal(rgb[:k], ws.rgb[:k])
    w = ws.w_sel[:k].double()
    assert float(w.min()) > 1e-4  # only masked-in samples were compacted (renderer_semantics.py:249-250)
    f64 = dict(dtype=torch.float64, device=DEV)
    image_ref = torch.zeros(n, 3, **f64).index_add_(0, ray, w[:, None] * rgb[:k].double())
    prob = torch.softmax(logits[:k, :c].double(), dim=-1)
    sem_ref = torch.zeros(n, c, **f64).index_add_(0, ray, w[:, None] * prob)
    depth_ref = torch.zeros(n, **f64).index_add_(0, ray, w * ws.z_sel[:k].double()) / dn.double()
    for name, got, want in (("image", image, image_ref), ("semantics", sem, sem_ref), ("depth", ws.depth, depth_ref),
                            ("image(chain)", ws.image, image_ref), ("semantics(chain)", ws.semantics, sem_ref)):
        torch.testing.assert_close(got.double(), want, rtol=1e-5, atol=1e-7, msg=lambda m, name=name: f"{name}: {m}")
    # backward: dL/dw_k = g_image . rgb_k + g_depth * z_k / norm  (semantic weights are detached, :270)
    gi, gd, gs = (torch.randn(n, 3, generator=g).to(DEV), torch.randn(n, generator=g).to(DEV),
                  torch.randn(n, c, generator=g).to(DEV))
    gc, gsw = torch.zeros(ops.COLOR_PARAMS, device=DEV), torch.zeros(ops.SEM_PARAMS, device=DEV)
    ops.heads_bwd(ws.sel, ws.ray_off, n, tc + tf, ws.k_max, d, ws.h, net.color_net.half_params(),
                  net.semantics_net.half_params(), c, ws.rgb, ws.hc1, ws.hc2, ws.hs, ws.w_sel, ws.z_sel, gi, gd, gs, dn,
                  128.0, ws.dh, ws.d_w_sel, gc, gsw)
    torch.cuda.synchronize()
    dw_ref = (gi.double()[ray] * rgb[:k].double()).sum(-1) + gd.double()[ray] * ws.z_sel[:k].double() / dn.double()[ray]
    torch.testing.assert_close(ws.d_w_sel[:k].double(), dw_ref, rtol=1e-5, atol=1e-5 * float(dw_ref.abs().max()) * 0.1)
