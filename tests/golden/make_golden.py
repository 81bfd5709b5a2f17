"""Generate the golden fixtures in this directory from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What is executed verbatim from the reference:
    nr4seg/nerf/renderer_semantics.py   (SemanticNeRFRenderer.run / render, sample_pdf)
    nr4seg/nerf/network_tcnn_semantics.py (SemanticNeRFNetwork.density/color/semantics)
    nr4seg/nerf/activation.py            (trunc_exp)

What has to be stubbed, and how:
    tinycudann   -> oracle/tcnn_spec.py (tcnn is absent; that arithmetic is "parity unpinned")
    trimesh      -> empty module (debug plotting only, renderer_semantics.py:49-58)
    nr4seg.nerf.raymarching -> near_far_from_aabb from oracle/live_path.py (the reference's
                    is a CUDA kernel; it is pinned on the GPU box against oracle/_ref instead)
    torch.rand   -> a queue of pre-drawn tensors, so the two random draws of run()
                    (renderer_semantics.py:166 and :28) are reproducible inputs
    trunc_exp    -> explicit .float() on its input: torch.cuda.amp.custom_fwd(cast_inputs=fp32)
                    only casts under CUDA autocast, which is how the reference always runs it
                    (joint_train_lightning_net.py:167,225)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import live_path, tcnn_spec as spec  # noqa: E402


# ----------------------------------------------------------------------------- stubs
class _Encoding(torch.nn.Module):
    def __init__(self, n_input_dims, encoding_config, **_):
        super().__init__()
        self.otype = encoding_config["otype"]
        if self.otype == "HashGrid":
            assert encoding_config["n_levels"] == 16 and encoding_config["log2_hashmap_size"] == 19
            self.n_output_dims = 32
            self.bound = STATE["bound"]
            self.table = spec.level_table(self.bound)
            assert abs(encoding_config["per_level_scale"] - spec.per_level_scale(self.bound)) < 1e-12
            n = self.table["total"] * spec.N_FEATURES
            self.params = torch.nn.Parameter(
                spec.splitmix_uniform(n, STATE["seed"], -STATE["hash_amp"], STATE["hash_amp"]))
        else:
            assert self.otype == "SphericalHarmonics" and encoding_config["degree"] == 4
            self.n_output_dims = 16
            self.params = torch.nn.Parameter(torch.zeros(0))

    def forward(self, x):
        if self.otype == "HashGrid":
            return spec.hashgrid_forward(x.float(), self.params, self.table).half()
        return spec.sh4_forward(x.float()).half()


class _Network(torch.nn.Module):
    _count = 0

    def __init__(self, n_input_dims, n_output_dims, network_config, **_):
        super().__init__()
        assert network_config["otype"] == "FullyFusedMLP" and network_config["activation"] == "ReLU"
        self.n_in, self.n_out = n_input_dims, n_output_dims
        self.dims = spec.mlp_dims(n_input_dims, n_output_dims, network_config["n_hidden_layers"],
                                  network_config["n_neurons"])
        _Network._count += 1
        self.params = torch.nn.Parameter(spec.xavier_mlp_init(self.dims, STATE["seed"] + _Network._count))

    def forward(self, x):
        return spec.mlp_forward(x.float(), self.params, self.dims, self.n_in, self.n_out).half()


STATE = {"bound": 4.0, "seed": 1337, "hash_amp": 0.5}
LOSS_SCALE = 1024.0


def install_stubs():
    sys.modules["trimesh"] = types.ModuleType("trimesh")
    tcnn = types.ModuleType("tinycudann")
    tcnn.Encoding, tcnn.Network = _Encoding, _Network
    sys.modules["tinycudann"] = tcnn
    pkg = types.ModuleType("nr4seg.nerf.raymarching")
    rm = types.ModuleType("nr4seg.nerf.raymarching.raymarching")
    rm.near_far_from_aabb = lambda o, d, aabb, min_near=0.2: live_path.near_far(o, d, aabb, min_near)
    pkg.raymarching = rm
    sys.modules["nr4seg.nerf.raymarching"] = pkg
    sys.modules["nr4seg.nerf.raymarching.raymarching"] = rm


class RandQueue:
    """Context manager replacing torch.rand by a FIFO of prepared tensors."""

    def __init__(self, tensors):
        self.q = list(tensors)

    def __enter__(self):
        self.orig = torch.rand

        def fake(*shape, **kw):
            t = self.q.pop(0)
            want = list(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else list(shape)
            assert list(t.shape) == [int(s) for s in want], (t.shape, want)
            return t.clone()

        torch.rand = fake
        return self

    def __exit__(self, *a):
        torch.rand = self.orig
        assert not self.q, "unused injected random tensors"


def make_rays(n, seed, n_outside=0):
    g = torch.Generator().manual_seed(seed)
    o = (torch.rand(n, 3, generator=g) - 0.5) * 2.0  # cameras well inside bound=4
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    dn = 1.0 + 0.3 * torch.rand(n, 1, generator=g)
    if n_outside:
        o[:n_outside] = torch.tensor([9.0, 9.5, 10.0])
        d[:n_outside] = torch.nn.functional.normalize(torch.tensor([[1.0, 0.2, 0.1]]), dim=-1)
    return o.unsqueeze(0), d.unsqueeze(0), dn.unsqueeze(0)


def sub_sample_grad(g: torch.Tensor, keep=4096):
    nz = torch.nonzero(g).view(-1)
    if nz.numel() > keep:
        nz = nz[:: nz.numel() // keep][:keep]
    return nz.numpy().astype(np.int64), g[nz].numpy(), float(g.double().sum()), float(g.double().abs().sum())


def case_network(name, *, n, steps, up, perturb, train, staged, max_ray_batch, n_outside, seed, backward):
    from nr4seg.nerf import network_tcnn_semantics as ref_net

    from nr4seg.nerf import activation as ref_act

    ref_net.trunc_exp = lambda x: ref_act.trunc_exp(x.float())
    _Network._count = 0
    model = ref_net.SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                                        num_semantic_classes=40)
    model.train(train)
    o, d, dn = make_rays(n, seed, n_outside)
    g = torch.Generator().manual_seed(seed + 99)
    t_rand = torch.rand(n, steps, generator=g)
    u = torch.rand(n, up, generator=g)
    # the chunk loop draws one u per chunk (and one t_rand when perturb)
    queue = []
    if staged:
        for head in range(0, n, max_ray_batch):
            tail = min(head + max_ray_batch, n)
            if perturb:
                queue.append(t_rand[head:tail])
            queue.append(u[head:tail])
    else:
        if perturb:
            queue.append(t_rand)
        queue.append(u)
    kw = dict(num_steps=steps, upsample_steps=up)
    with RandQueue(queue):
        out = model.render(o, d, direction_norms=dn, staged=staged, max_ray_batch=max_ray_batch,
                           bg_color=None, perturb=perturb, **kw)
    blob = dict(rays_o=o.numpy(), rays_d=d.numpy(), direction_norms=dn.numpy(), t_rand=t_rand.numpy(),
                u=u.numpy(), depth=out["depth"].detach().numpy(), image=out["image"].detach().numpy(),
                semantics=out["semantics"].detach().numpy(),
                cfg=np.array([n, steps, up, int(perturb), int(train), int(staged), max_ray_batch, 40,
                              STATE["seed"]], dtype=np.int64),
                hash_amp=np.float32(STATE["hash_amp"]))
    if backward:
        gi = torch.randn(out["image"].shape, generator=g)
        gd = torch.randn(out["depth"].shape, generator=g)
        gs = torch.randn(out["semantics"].shape, generator=g)
        loss = (out["image"] * gi).sum() + (out["depth"] * gd).sum() + (out["semantics"] * gs).sum()
        # gradients cross the fp16 tensors of the reference code in fp16; scale like the reference's
        # GradScaler does (joint_train_lightning_net.py:46,509) so small ones do not flush to zero
        (loss * LOSS_SCALE).backward()
        for p in model.parameters():
            if p.grad is not None:
                p.grad /= LOSS_SCALE
        idx, val, s, a = sub_sample_grad(model.encoder.params.grad)
        blob.update(g_image=gi.numpy(), g_depth=gd.numpy(), g_semantics=gs.numpy(),
                    grad_sigma_net=model.sigma_net.params.grad.numpy(),
                    grad_color_net=model.color_net.params.grad.numpy(),
                    grad_semantics_net=model.semantics_net.params.grad.numpy(),
                    grad_hash_idx=idx, grad_hash_val=val, grad_hash_sum=np.float64(s),
                    grad_hash_abs=np.float64(a))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in blob.items()})


def case_composite(name, *, n, steps, seed):
    """Config-1 shaped: synthetic heads isolate weights/masks/compositing."""
    from nr4seg.nerf.renderer_semantics import SemanticNeRFRenderer

    g = torch.Generator().manual_seed(seed)
    sigma = (50 * torch.rand(n * steps, generator=g) ** 4).requires_grad_()
    rgb = torch.rand(n * steps, 3, generator=g).requires_grad_()
    prob = torch.softmax(torch.randn(n * steps, 40, generator=g), dim=-1).requires_grad_()

    class Synth(SemanticNeRFRenderer):
        def density(self, x):
            return {"sigma": sigma, "geo_feat": torch.zeros(x.shape[0], 1)}

        def color(self, x, d, mask=None, **kw):
            return rgb * mask.unsqueeze(1)

        def semantics(self, x, d, mask=None, **kw):
            return prob * mask.unsqueeze(1)

    model = Synth(bound=4, cuda_ray=False, density_scale=1, num_semantic_classes=40)
    o = torch.zeros(1, n, 3)
    d = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1)
    dn = 1.0 + 0.2 * torch.rand(1, n, 1, generator=g)
    out = model.run(o, d, dn, num_steps=steps, upsample_steps=0, perturb=False)
    gi = torch.randn(out["image"].shape, generator=g)
    gd = torch.randn(out["depth"].shape, generator=g)
    gs = torch.randn(out["semantics"].shape, generator=g)
    ((out["image"] * gi).sum() + (out["depth"] * gd).sum() + (out["semantics"] * gs).sum()).backward()
    blob = dict(sigma=sigma.detach().numpy(), rgb=rgb.detach().numpy(), prob=prob.detach().numpy(),
                rays_d=d.numpy(), direction_norms=dn.numpy(), depth=out["depth"].detach().numpy(),
                image=out["image"].detach().numpy(), semantics=out["semantics"].detach().numpy(),
                g_image=gi.numpy(), g_depth=gd.numpy(), g_semantics=gs.numpy(),
                grad_sigma=sigma.grad.numpy(), grad_rgb=rgb.grad.numpy(), grad_prob=prob.grad.numpy(),
                cfg=np.array([n, steps, 40], dtype=np.int64))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
    print(name, "ok")


def case_sample_pdf(name, *, n, bins, n_samples, seed):
    from nr4seg.nerf.renderer_semantics import sample_pdf

    g = torch.Generator().manual_seed(seed)
    z = torch.sort(torch.rand(n, bins, generator=g) * 5 + 0.2, dim=1).values
    w = torch.rand(n, bins - 1, generator=g) ** 6
    w[0] = 0  # all-zero weights -> uniform pdf from the +1e-5
    w[1, : bins // 2] = 0
    u = torch.rand(n, n_samples, generator=g)
    with RandQueue([u]):
        out = sample_pdf(z, w, n_samples, det=False)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), bins=z.numpy(), weights=w.numpy(), u=u.numpy(),
                        samples=out.numpy())
    print(name, "ok")


def case_frontend(name, *, height, width, n, seed):
    """Rows f2 / f3 / f4: ray generation, ground-truth gather, pseudo-label epilogue, pose conventions.
    Executed from the reference: nr4seg/dataset/ngp_utils.py (get_rays, nerf_matrix_to_ngp; loaded by file path - the
    package __init__ pulls in imageio / cv2).  get_rays_train (joint_train_lightning_net.py:109-151) is get_rays
    followed by a gather of the sampled pixels, and the epilogue (:246-250, :755-768) and the Slerp interpolation
    (scannet_ngp_joint.py:229-262) are methods of classes that need Lightning / OpenCV: those few lines are replayed
    here with the same torch / scipy calls."""
    import importlib.util

    from scipy.spatial.transform import Rotation, Slerp

    spec_ = importlib.util.spec_from_file_location("ref_ngp_utils", "/root/reference/nr4seg/dataset/ngp_utils.py")
    ngp_utils = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(ngp_utils)

    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    nerf_poses = []
    for _ in range(6):
        p = np.eye(4, dtype=np.float32)
        p[:3, :3] = Rotation.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        p[:3, 3] = rng.normal(size=3)
        nerf_poses.append(p)
    ngp_poses = np.stack([ngp_utils.nerf_matrix_to_ngp(p) for p in nerf_poses])
    intrinsics = np.array([0.9 * width, 0.95 * width, width / 2 - 0.3, height / 2 + 0.7])
    rays = ngp_utils.get_rays(torch.from_numpy(ngp_poses[:2]), intrinsics, height, width)
    inds = torch.randint(0, height * width, size=[n], generator=g)  # may duplicate, as :142
    # ground truth planes and the three gathers of forward_nerf_train (:180-187)
    image = torch.rand(1, 3, height, width, generator=g).half()
    labels = torch.randint(0, 40, (1, height, width), generator=g)
    depth = torch.rand(1, height, width, generator=g) * 4
    b_inds = inds.expand([1, n])
    gt_rgb = torch.gather(image.reshape(1, 3, -1).permute(0, 2, 1), 1, torch.stack(3 * [b_inds], -1))
    gt_labels = torch.gather(labels.view(1, -1), 1, b_inds)
    gt_depth = torch.gather(depth.view(1, -1), 1, b_inds)
    # pseudo-label epilogue (:246-250, :755-768)
    semantics = torch.rand(n, 40, generator=g) ** 3
    semantics[::7] = 0  # pixels without semantic mass
    rgb = torch.rand(n, 3, generator=g)
    sem = semantics.clone()
    invalid = torch.sum(sem, dim=-1) == 0
    sem[invalid] = 1
    sem = sem / torch.sum(sem, dim=-1, keepdim=True)
    label_u8 = (torch.argmax(sem, dim=-1) + 1).numpy().astype(np.uint8)
    rgb_u8 = (rgb.numpy() * 255).astype(np.uint8)
    # novel view points (scannet_ngp_joint.py:229-262)
    ring = nerf_poses + [nerf_poses[0]]
    slerp = Slerp(times=[*range(len(ring))], rotations=Rotation.from_matrix([p[:3, :3] for p in ring]))
    rots = slerp(times=[0.5 + i for i in range(len(ring) - 1)]).as_matrix()
    novel = []
    for i in range(len(ring) - 1):
        q = np.eye(4)
        q[:3, :3] = rots[i]
        q[:3, 3] = (ring[i][:3, 3] + ring[i + 1][:3, 3]) / 2.0
        novel.append(q)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), height=height, width=width, intrinsics=intrinsics,
        nerf_poses=np.stack(nerf_poses), ngp_poses=ngp_poses, rays_o=rays["rays_o"].numpy(),
        rays_d=rays["rays_d"].numpy(), direction_norms=rays["direction_norms"].numpy(), inds=inds.numpy(),
        image_h=image[0].numpy(), labels=labels[0].numpy(), depth=depth[0].numpy(), gt_rgb=gt_rgb[0].numpy(),
        gt_labels=gt_labels[0].numpy(), gt_depth=gt_depth[0].numpy(), semantics=semantics.numpy(), rgb=rgb.numpy(),
        label_u8=label_u8, rgb_u8=rgb_u8, novel_poses=np.stack(novel))
    print(name, "rays", tuple(rays["rays_d"].shape), "novel", len(novel))


def install_lightning_stubs():
    """What joint_train_lightning_net.py imports besides torch / torchvision / cv2 / PIL, none of it on the NeRF path:
    pytorch_lightning (absent; LightningModule -> nn.Module), the DeepLab wrapper, the metric and visualizer helpers
    (Lightning / matplotlib / imageio dependent).  nr4seg.dataset.ngp_utils is the reference's own file, loaded by
    path because the package __init__ imports the dataset classes."""
    import importlib.util

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        current_epoch = 0

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl
    for name, attrs in (("nr4seg.network", ["DeepLabV3"]), ("nr4seg.utils", []), ("nr4seg.utils.metrics", ["SemanticsMeter"]),
                        ("nr4seg.visualizer", ["Visualizer"]), ("nr4seg.dataset", [])):
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, type(a, (), {"__init__": lambda self, *args, **kw: None}))
        sys.modules[name] = mod
    spec_ = importlib.util.spec_from_file_location("nr4seg.dataset.ngp_utils",
                                                   "/root/reference/nr4seg/dataset/ngp_utils.py")
    ngp_utils = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(ngp_utils)
    sys.modules["nr4seg.dataset.ngp_utils"] = ngp_utils
    return ngp_utils


def case_lightning(name, *, height, width, seed):
    """The CONSUMER side of the boundary: the reference's unmodified JointTrainLightningNet.forward_nerf_train and
    forward_nerf_test (joint_train_lightning_net.py:167-257) drive the reference network (tcnn stubbed as above) on a
    small view; the batch, the pixels they sampled and what they returned are frozen.  tests/test_gpu_boundary.py
    replays the same calls on the drop-in module under CUDA autocast."""
    ngp_utils = install_lightning_stubs()
    import importlib.util

    # the module file itself, unmodified; loaded by path because nr4seg/lightning/__init__.py imports the data modules
    spec_pl = importlib.util.spec_from_file_location(
        "ref_joint_train_lightning_net", "/root/reference/nr4seg/lightning/joint_train_lightning_net.py")
    ref_pl = importlib.util.module_from_spec(spec_pl)
    spec_pl.loader.exec_module(ref_pl)
    from nr4seg.nerf import activation as ref_act
    from nr4seg.nerf import network_tcnn_semantics as ref_net

    ref_net.trunc_exp = lambda x: ref_act.trunc_exp(x.float())
    _Network._count = 0
    nerf = ref_net.SemanticNeRFNetwork(encoding="hashgrid", bound=4, cuda_ray=False, density_scale=1,
                                       num_semantic_classes=40)
    mod = ref_pl.JointTrainLightningNet.__new__(ref_pl.JointTrainLightningNet)
    torch.nn.Module.__init__(mod)
    mod.num_classes = 40
    mod.nerf_model = nerf
    mod.criterion_nerf_rgb = torch.nn.MSELoss(reduction="none")
    mod.criterion_nerf_semantics = torch.nn.NLLLoss(ignore_index=-1, reduction="none")
    mod.criterion_nerf_depth = torch.nn.L1Loss(reduction="none")
    mod._default_H = mod._default_W = None

    g = torch.Generator().manual_seed(seed)
    n_pix = height * width
    # camera inside the box, looking around: nerf_matrix_to_ngp convention is whatever the caller hands over
    from scipy.spatial.transform import Rotation

    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = Rotation.from_euler("xyz", [0.3, -0.5, 0.2]).as_matrix()
    pose[:3, 3] = [0.4, -0.3, 0.2]
    intr = np.array([0.9 * width, 0.95 * width, width / 2 - 0.3, height / 2 + 0.7], dtype=np.float64)
    img = torch.rand(1, 3, height, width, generator=g)
    depth = torch.rand(1, height, width, generator=g) * 3
    depth[0, ::5, ::3] = 0  # invalid depth readings
    seg = torch.randint(0, 40, (1, height, width), generator=g)
    uom = 0.6
    batch = {"pose": torch.from_numpy(pose)[None], "intrinsics": [tuple(float(v) for v in intr)], "H": [height],
             "W": [width], "img_fp16": img.half().float(), "img": img, "depth": depth, "one_m_to_scene_uom": [uom]}
    # (img_fp16 is handed over as fp32 HOLDING fp16 values: under CUDA autocast mse_loss runs in fp32 on the promoted
    # fp16 pixels, and the CUDA-autocast decorators of the reference do nothing on this CPU-only build machine)
    steps = up = 256  # the defaults of render() -> run() (renderer_semantics.py:127-128)
    n_rays = min(4096, n_pix)
    t_rand = spec.splitmix_uniform(n_rays * steps, 901, 0.0, 1.0).view(n_rays, steps)
    u_train = spec.splitmix_uniform(n_rays * up, 902, 0.0, 1.0).view(n_rays, up)
    u_test = spec.splitmix_uniform(n_pix * up, 903, 0.0, 1.0).view(n_pix, up)

    captured = {}
    orig_rays = mod.get_rays_train

    def capture(batch_, bs_, N=4096):
        out = orig_rays(batch_, bs_, N)
        captured["inds"] = out[3].clone()
        return out

    mod.get_rays_train = capture
    torch.manual_seed(seed)
    nerf.train()
    with RandQueue([t_rand, u_train]):
        l_color, l_sem, l_depth = mod.forward_nerf_train(batch, {"seg_semantics": seg}, 0)
    total = l_color + l_sem * 0.04 + l_depth * 0.1  # training_step_nerf, :503-507
    (total * LOSS_SCALE).backward()
    grads = {k: (getattr(nerf, k).params.grad / LOSS_SCALE) for k in ("sigma_net", "color_net", "semantics_net")}
    # full frame, the pseudo-label pass (:225-257); rays as the DataLoader builds them (ngp_utils.get_rays)
    rays = ngp_utils.get_rays(torch.from_numpy(pose)[None], intr, height, width)
    tbatch = {"rays_o": rays["rays_o"], "rays_d": rays["rays_d"], "direction_norms": rays["direction_norms"],
              "viewpoint_is_novel": [False], "img": img}
    nerf.eval()
    with torch.no_grad(), RandQueue([u_test]):
        test_out = mod.forward_nerf_test(tbatch)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), height=height, width=width, pose=pose, intrinsics=intr,
        img=img.numpy(), depth=depth.numpy(), seg=seg.numpy(), uom=np.float32(uom), inds=captured["inds"].numpy(),
        losses=np.array([float(l_color), float(l_sem), float(l_depth), float(total)], dtype=np.float64),
        grad_sigma_net=grads["sigma_net"].numpy(), grad_color_net=grads["color_net"].numpy(),
        grad_semantics_net=grads["semantics_net"].numpy(),
        test_rays_o=rays["rays_o"].numpy(), test_rays_d=rays["rays_d"].numpy(),
        test_norms=rays["direction_norms"].numpy(), nerf_rgb=test_out["nerf_rgb"].numpy(),
        nerf_semantics=test_out["nerf_semantics"].numpy().astype(np.int64),
        nerf_semantics_raw=test_out["nerf_semantics_raw"].numpy().astype(np.float32),
        cfg=np.array([steps, up, 40, STATE["seed"], 901, 902, 903], dtype=np.int64), hash_amp=np.float32(STATE["hash_amp"]))
    print(name, "losses", float(l_color), float(l_sem), float(l_depth), "rays", n_rays, "frame", test_out["nerf_rgb"].shape)



def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    install_stubs()
    case_sample_pdf("sample_pdf", n=16, bins=33, n_samples=48, seed=5)
    case_composite("composite_cfg1_small", n=64, steps=128, seed=1234)
    case_network("train_small", n=96, steps=24, up=24, perturb=True, train=True, staged=False,
                 max_ray_batch=4096, n_outside=0, seed=11, backward=True)
    case_network("infer_staged", n=70, steps=16, up=16, perturb=False, train=False, staged=True,
                 max_ray_batch=32, n_outside=2, seed=12, backward=False)
    case_frontend("frontend", height=24, width=40, n=300, seed=21)
    case_lightning("lightning_step", height=20, width=32, seed=31)


if __name__ == "__main__":
    main()
