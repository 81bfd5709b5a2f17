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
