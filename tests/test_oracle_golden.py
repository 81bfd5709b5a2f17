"""The CPU oracle (oracle/live_path.py + oracle/tcnn_spec.py) against fixtures frozen from the
reference's own modules by tests/golden/make_golden.py.  CPU only."""
import os

import numpy as np
import torch

from oracle import live_path, tcnn_spec as spec


def _load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_level_table_matches_survey():
    t = spec.level_table(4)
    assert list(t["res"]) == [16, 25, 37, 56, 85, 128, 195, 295, 446, 676, 1024, 1553, 2353, 3566, 5405, 8192]
    assert t["total"] == 6537456 and t["total"] * 2 == 13074912
    assert list(t["hashed"]) == [False] * 4 + [True] * 12


def test_sample_pdf_golden(golden_dir):
    g = _load(golden_dir, "sample_pdf")
    out = live_path.sample_pdf(_t(g["bins"]), _t(g["weights"]), g["u"].shape[1], u=_t(g["u"]))
    np.testing.assert_allclose(out.numpy(), g["samples"], rtol=1e-6, atol=1e-6)


def test_composite_golden(golden_dir):
    g = _load(golden_dir, "composite_cfg1_small")
    n, steps, c = [int(v) for v in g["cfg"]]
    sigma = _t(g["sigma"]).requires_grad_()
    rgb = _t(g["rgb"]).requires_grad_()
    prob = _t(g["prob"]).requires_grad_()

    class Synth:
        bound = 4.0
        num_semantic_classes = c

        def density(self, x):
            return {"sigma": sigma, "geo_feat": torch.zeros(x.shape[0], 1)}

        def color(self, x, d, mask=None, **kw):
            return rgb * mask.unsqueeze(1)

        def semantics(self, x, d, mask=None, **kw):
            return prob * mask.unsqueeze(1)

    out = live_path.run(Synth(), torch.zeros(1, n, 3), _t(g["rays_d"]), _t(g["direction_norms"]),
                        num_steps=steps, upsample_steps=0, perturb=False)
    for k in ("depth", "image", "semantics"):
        np.testing.assert_allclose(out[k].detach().numpy(), g[k], rtol=1e-5, atol=1e-6)
    loss = (out["image"] * _t(g["g_image"])).sum() + (out["depth"] * _t(g["g_depth"])).sum() \
        + (out["semantics"] * _t(g["g_semantics"])).sum()
    loss.backward()
    np.testing.assert_allclose(sigma.grad.numpy(), g["grad_sigma"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rgb.grad.numpy(), g["grad_rgb"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(prob.grad.numpy(), g["grad_prob"], rtol=1e-5, atol=1e-7)


def _heads(g):
    seed = int(g["cfg"][8])
    return live_path.OracleHeads(bound=4, num_semantic_classes=int(g["cfg"][7]), seed=seed,
                                 hash_amp=float(g["hash_amp"]))


def test_train_small_golden(golden_dir):
    g = _load(golden_dir, "train_small")
    n, steps, up = [int(v) for v in g["cfg"][:3]]
    heads = _heads(g)
    out = live_path.run(heads, _t(g["rays_o"]), _t(g["rays_d"]), _t(g["direction_norms"]), num_steps=steps,
                        upsample_steps=up, perturb=True, t_rand=_t(g["t_rand"]), u=_t(g["u"]))
    for k in ("depth", "image", "semantics"):
        np.testing.assert_allclose(out[k].detach().numpy(), g[k], rtol=2e-5, atol=2e-6, err_msg=k)
    loss = (out["image"] * _t(g["g_image"])).sum() + (out["depth"] * _t(g["g_depth"])).sum() \
        + (out["semantics"] * _t(g["g_semantics"])).sum()
    loss.backward()
    for name, p in (("sigma_net", heads.sigma_net), ("color_net", heads.color_net),
                    ("semantics_net", heads.semantics_net)):
        ref = g["grad_" + name]
        # the fixture's gradients crossed fp16 tensors (2^-11 rounding per hop); the oracle's are fp32
        np.testing.assert_allclose(p.grad.numpy(), ref, rtol=5e-3, atol=2e-3 * np.abs(ref).max(), err_msg=name)
    gh = heads.encoder.grad
    np.testing.assert_allclose(gh[_t(g["grad_hash_idx"])].numpy(), g["grad_hash_val"], rtol=5e-3,
                               atol=2e-3 * np.abs(g["grad_hash_val"]).max())
    assert abs(float(gh.double().abs().sum()) - float(g["grad_hash_abs"])) < 2e-3 * float(g["grad_hash_abs"])


def test_infer_staged_golden(golden_dir):
    g = _load(golden_dir, "infer_staged")
    n, steps, up = [int(v) for v in g["cfg"][:3]]
    heads = _heads(g)
    with torch.no_grad():
        out = live_path.render(heads, _t(g["rays_o"]), _t(g["rays_d"]), _t(g["direction_norms"]), staged=True,
                               max_ray_batch=int(g["cfg"][6]), num_steps=steps, upsample_steps=up,
                               perturb=False, u=_t(g["u"]))
    for k in ("depth", "image", "semantics"):
        np.testing.assert_allclose(out[k].numpy(), g[k], rtol=2e-5, atol=2e-6, err_msg=k)


def test_pcg32_known_answer():
    """pcg32.h:44-116 against the published PCG32 demo stream (seed 42, stream 54)."""
    from oracle import raymarch

    u, f = raymarch.pcg32_stream(42, 54, 6)
    assert [int(x) for x in u] == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]
    assert ((f >= 0) & (f < 1)).all() and abs(float(f[0]) - ((0xa15c02b7 >> 9) / 2 ** 23)) < 1e-7


def test_c_oracle_march_properties():
    """C restatement of the marcher: fully occupied grid -> every step is dt = clamp(t*dt_gamma), samples stay inside
    the box, counts are capped at 1024; empty grid -> no samples."""
    from oracle import raymarch

    g = torch.Generator().manual_seed(0)
    n = 50
    o = ((torch.rand(n, 3, generator=g) - 0.5)).numpy()
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).numpy()
    aabb = np.array([-4, -4, -4, 4, 4, 4], dtype=np.float32)
    nears, fars = raymarch.near_far(o, d, aabb)
    ref_n, ref_f = live_path.near_far(torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(aabb))
    assert np.array_equal(nears, ref_n.numpy()) and np.array_equal(fars, ref_f.numpy())
    full = np.ones((3, 128, 128, 128), dtype=np.float32)
    xyz, dirs, deltas, rays, cnt = raymarch.march_rays_train(o, d, full, 1.0, 4.0, 1 / 128, nears, fars, n * 1024, 0)
    assert cnt[1] == n and cnt[0] == rays[:, 2].sum() and (rays[:, 2] <= 1024).all() and (rays[:, 2] > 10).all()
    tot = cnt[0]
    assert np.abs(xyz[:tot]).max() <= 4.0 and (deltas[:tot, 0] >= 2 * 1.73205080757 / 1024 - 1e-7).all()
    assert (deltas[:tot, 0] <= 2 * 4.0 / 128 + 1e-7).all()
    np.testing.assert_allclose(deltas[:tot, 0], deltas[:tot, 1], rtol=1e-4)  # no skipped voxels: the two deltas agree
    empty = np.zeros_like(full)
    _, _, _, rays0, cnt0 = raymarch.march_rays_train(o, d, empty, 1.0, 4.0, 1 / 128, nears, fars, n * 1024, 0)
    assert cnt0[0] == 0 and (rays0[:, 2] == 0).all()


def test_oracle_reproduces_the_unmodified_lightning_step(golden_dir):
    """tests/golden/lightning_step.npz: losses and pseudo labels the reference's own forward_nerf_train /
    forward_nerf_test (joint_train_lightning_net.py:167-257) returned; the oracle (ray generation, live path, losses)
    must reproduce them -- this pins oracle/frontend.py + live_path.py + losses.py to the consumer of the path."""
    from oracle import frontend
    from oracle.losses import nerf_losses

    g = {k: v for k, v in np.load(os.path.join(golden_dir, "lightning_step.npz")).items()}
    h, w = int(g["height"]), int(g["width"])
    steps, up, c, seed = (int(v) for v in g["cfg"][:4])
    s1, s2, s3 = (int(v) for v in g["cfg"][4:7])
    heads = live_path.OracleHeads(bound=4, num_semantic_classes=c, seed=seed, hash_amp=float(g["hash_amp"]))
    inds = torch.from_numpy(g["inds"])[0]
    n = inds.numel()
    o, d, dn = frontend.get_rays(g["pose"], g["intrinsics"], h, w, inds=inds)
    img16 = torch.from_numpy(g["img"]).half().float()[0]
    gt_rgb, labels, gt_depth = frontend.gather_gt(img16, g["seg"][0], g["depth"][0], inds)
    t_rand = spec.splitmix_uniform(n * steps, s1, 0.0, 1.0).view(n, steps)
    u = spec.splitmix_uniform(n * up, s2, 0.0, 1.0).view(n, up)
    out = live_path.run(heads, o[None], d[None], dn[None, :, None], num_steps=steps, upsample_steps=up, perturb=True,
                        t_rand=t_rand, u=u)
    total, parts = nerf_losses(out, gt_rgb[None], labels[None], gt_depth[None], float(g["uom"]))
    got = np.array([float(p) for p in parts] + [float(total)])
    np.testing.assert_allclose(got, g["losses"], rtol=2e-5, atol=1e-7)
    # full-frame pseudo labels
    u_test = spec.splitmix_uniform(h * w * up, s3, 0.0, 1.0).view(h * w, up)
    with torch.no_grad():
        full = live_path.render(heads, torch.from_numpy(g["test_rays_o"]), torch.from_numpy(g["test_rays_d"]),
                                torch.from_numpy(g["test_norms"]), staged=True, num_steps=steps, upsample_steps=up,
                                perturb=False, u=u_test)
    np.testing.assert_allclose(full["image"].reshape(1, h, w, 3).permute(0, 3, 1, 2).numpy(), g["nerf_rgb"], rtol=2e-5,
                               atol=1e-6)
    label_u8, _ = frontend.label_epilogue(full["semantics"][0].numpy(), full["image"][0].numpy())
    assert (label_u8.reshape(h, w).astype(np.int64) - 1 == g["nerf_semantics"][0]).mean() > 0.999


def test_sh4_spec_is_the_published_real_spherical_harmonics():
    """Row a7 has no tcnn build to pin against; what CAN be pinned independently is the mathematics: the 16 polynomials of
    oracle/tcnn_spec.sh4_forward (network_tcnn_semantics.py:64-70: SphericalHarmonics, degree 4) must be the real
    spherical harmonics of degree 0..3 in the order l^2 + l + m with the Condon-Shortley phase kept, evaluated here from
    scipy's COMPLEX harmonics (sqrt(2) Re Y_l^m for m > 0, sqrt(2) Im Y_l^|m| for m < 0), not from polynomial constants."""
    from scipy.special import sph_harm_y

    rng = np.random.default_rng(0)
    d = rng.normal(size=(2000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    theta, phi = np.arccos(np.clip(d[:, 2], -1, 1)), np.arctan2(d[:, 1], d[:, 0])
    cols = []
    for l in range(4):
        for m in range(-l, l + 1):
            y = sph_harm_y(l, abs(m), theta, phi)
            cols.append(y.real if m == 0 else np.sqrt(2) * (y.real if m > 0 else y.imag))
    want = np.stack(cols, axis=1)
    got = spec.sh4_forward(_t(((d + 1) / 2).astype(np.float32))).numpy()
    # fp16 output: half an ulp of values up to ~0.75, plus the fp32 rounding of (d + 1) / 2 -> 2 x - 1
    np.testing.assert_allclose(got, want, rtol=0, atol=2.6e-4)


def test_mlp_spec_against_plain_half_precision_linear_layers():
    """Row a6: tcnn_spec.mlp_forward (weights [out, in] per layer, fp32 accumulate, ReLU, fp16 between layers, constant-1
    padding column) written a second time with torch.nn.functional.linear on float64 copies of the fp16 operands."""
    g = torch.Generator().manual_seed(5)
    dims, n_in, n_out = (32, 64, 64, 16), 31, 3
    n_par = sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    params = (torch.rand(n_par, generator=g) - 0.5) * 0.5
    x = (torch.randn(257, n_in, generator=g)).half().float()
    got = spec.mlp_forward(x, params, dims, n_in, n_out)
    h = torch.cat([x, torch.ones(257, 1)], dim=1).double()
    off = 0
    for i, (fi, fo) in enumerate(zip(dims[:-1], dims[1:])):
        w = params[off:off + fi * fo].half().double().view(fo, fi)
        off += fi * fo
        h = torch.nn.functional.linear(h, w)
        if i < len(dims) - 2:
            h = torch.relu(h)
        h = h.float().half().double()
    # the float64 product rounds once; the fp32 accumulation of the spec may land on the neighbouring fp16 value
    np.testing.assert_allclose(got.numpy(), h[:, :n_out].float().numpy(), rtol=2e-3, atol=1e-4)
