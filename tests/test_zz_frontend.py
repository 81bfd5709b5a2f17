"""Rows f2 / f3 / f4 of SURVEY.md section 8: ray generation, ground-truth gather, pseudo-label epilogue, pose formats.
CPU part: the oracle restatement and the host-side pose module against the fixture generated from the reference's own
code (tests/golden/make_golden.py: case_frontend).  GPU part: the kernels against the fixture, through the C ABI."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import frontend as oracle_frontend
from ucsa_neural_rendering_b200 import poses

DEV = "cuda"


@pytest.fixture(scope="module")
def fx():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend.npz")
    return {k: v for k, v in np.load(path).items()}


# ------------------------------------------------------------------------------------------------ CPU
def test_pose_conventions_match_reference(fx):
    for nerf, ngp in zip(fx["nerf_poses"], fx["ngp_poses"]):
        out = poses.nerf_matrix_to_ngp(nerf)
        assert out.dtype == np.float32 and np.array_equal(out, ngp)


def test_novel_view_interpolation_matches_scipy_slerp(fx):
    mine = poses.interpolate_novel_poses(list(fx["nerf_poses"]))
    assert len(mine) == len(fx["novel_poses"])  # the ring is closed: last -> first
    for a, b in zip(mine, fx["novel_poses"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)
        np.testing.assert_allclose(a[:3, :3] @ a[:3, :3].T, np.eye(3), atol=1e-6)  # still a rotation
    # a half turn about an axis is the hard case of the log map
    r = np.diag([1.0, -1.0, -1.0])
    p0, p1 = np.eye(4), np.eye(4)
    p1[:3, :3] = r
    mid = poses.interpolate_novel_poses([p0, p1])[0][:3, :3]
    np.testing.assert_allclose(mid @ mid, r, atol=1e-6)


def test_transforms_json_reader(tmp_path, fx):
    frames = [{"file_path": f"color/{i}.jpg", "label_path": f"label/{i}.png",
               "transform_matrix": fx["nerf_poses"][i % 6].tolist()} for i in range(10)]
    info = {"h": 480, "w": 640, "fl_x": 577.0, "fl_y": 578.0, "cx": 319.5, "cy": 239.5, "one_m_to_scene_uom": 0.41,
            "frames": frames}
    with open(tmp_path / "transforms_train.json", "w") as fh:
        json.dump(info, fh)
    train = poses.load_transforms(str(tmp_path), "train")
    val = poses.load_transforms(str(tmp_path), "val")
    pred = poses.load_transforms(str(tmp_path), "predict")
    assert (len(train.poses), len(val.poses), len(pred.poses)) == (8, 2, 10)  # last 20 % held out
    assert train.height == 480 and train.width == 640 and train.one_m_to_scene_uom == 0.41
    assert list(train.intrinsics) == [577.0, 578.0, 319.5, 239.5]
    assert np.array_equal(train.poses[3], fx["ngp_poses"][3])
    assert train.depth_paths[0].endswith(os.path.join("depth", "0.png"))
    assert val.image_paths[0].endswith(os.path.join("color", "8.jpg"))


def test_oracle_frontend_matches_reference_fixture(fx):
    h, w = int(fx["height"]), int(fx["width"])
    for v in range(2):
        o, d, n = oracle_frontend.get_rays(fx["ngp_poses"][v], fx["intrinsics"], h, w)
        np.testing.assert_array_equal(o.numpy(), fx["rays_o"][v])
        np.testing.assert_allclose(d.numpy(), fx["rays_d"][v], rtol=0, atol=1e-7)
        np.testing.assert_array_equal(n.numpy(), fx["direction_norms"][v][:, 0])
    rgb, lab, dep = oracle_frontend.gather_gt(fx["image_h"], fx["labels"], fx["depth"], fx["inds"])
    assert np.array_equal(rgb.numpy(), fx["gt_rgb"]) and np.array_equal(lab.numpy(), fx["gt_labels"])
    assert np.array_equal(dep.numpy(), fx["gt_depth"])
    lab_u8, rgb_u8 = oracle_frontend.label_epilogue(fx["semantics"], fx["rgb"])
    assert np.array_equal(lab_u8, fx["label_u8"]) and np.array_equal(rgb_u8, fx["rgb_u8"])


def test_front_end_ops_refuse_cpu_tensors(fx):
    """no CPU fallback on the product path: the wrappers raise instead of computing on the host"""
    from ucsa_neural_rendering_b200 import ops
    from ucsa_neural_rendering_b200._lib import UcsaError

    with pytest.raises(UcsaError):
        ops.generate_rays(torch.eye(4), fx["intrinsics"], 4, 4)
    with pytest.raises(UcsaError):
        ops.label_epilogue(torch.rand(8, 40))
    with pytest.raises(UcsaError):
        ops.gather_gt(torch.rand(3, 4, 4).half(), torch.zeros(2, dtype=torch.int64))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_generate_rays_kernel(fx):
    from ucsa_neural_rendering_b200 import ops

    h, w = int(fx["height"]), int(fx["width"])
    for v in range(2):
        pose = torch.from_numpy(fx["ngp_poses"][v]).to(DEV)
        o, d, n = ops.generate_rays(pose, fx["intrinsics"], h, w)  # the whole view
        np.testing.assert_array_equal(o.cpu().numpy(), fx["rays_o"][v])
        # fp32 throughout; torch's CPU norm / matmul round sums of three terms in another order: one ulp
        np.testing.assert_allclose(n.cpu().numpy(), fx["direction_norms"][v][:, 0], rtol=1.2e-7, atol=0)
        np.testing.assert_allclose(d.cpu().numpy(), fx["rays_d"][v], rtol=0, atol=2.5e-7)
    inds = torch.from_numpy(fx["inds"]).to(DEV)  # sampled pixels, duplicates allowed
    o, d, n = ops.generate_rays(torch.from_numpy(fx["ngp_poses"][0]).to(DEV), fx["intrinsics"], h, w, inds=inds)
    np.testing.assert_allclose(d.cpu().numpy(), fx["rays_d"][0][fx["inds"]], rtol=0, atol=2.5e-7)
    np.testing.assert_allclose(n.cpu().numpy(), fx["direction_norms"][0][fx["inds"], 0], rtol=1.2e-7, atol=0)
    # the sampled form equals the whole-view form at the same pixels, bit for bit
    o_all, d_all, n_all = ops.generate_rays(torch.from_numpy(fx["ngp_poses"][0]).to(DEV), fx["intrinsics"], h, w)
    assert torch.equal(d, d_all[inds]) and torch.equal(n, n_all[inds]) and torch.equal(o, o_all[inds])


@pytest.mark.gpu
def test_gather_gt_kernel_is_exact(fx):
    from ucsa_neural_rendering_b200 import ops

    inds = torch.from_numpy(fx["inds"]).to(DEV)
    rgb, lab, dep = ops.gather_gt(torch.from_numpy(fx["image_h"]).to(DEV), inds,
                                  labels=torch.from_numpy(fx["labels"]).to(DEV),
                                  depth=torch.from_numpy(fx["depth"]).to(DEV))
    assert np.array_equal(rgb.cpu().numpy(), fx["gt_rgb"])
    assert np.array_equal(lab.cpu().numpy(), fx["gt_labels"]) and np.array_equal(dep.cpu().numpy(), fx["gt_depth"])
    rgb2, lab2, dep2 = ops.gather_gt(torch.from_numpy(fx["image_h"]).to(DEV), inds)
    assert lab2 is None and dep2 is None and torch.equal(rgb2, rgb)


@pytest.mark.gpu
def test_label_epilogue_kernel_is_exact(fx):
    from ucsa_neural_rendering_b200 import ops

    sem = torch.from_numpy(fx["semantics"]).to(DEV)
    rgb = torch.from_numpy(fx["rgb"]).to(DEV)
    lab_u8, rgb_u8 = ops.label_epilogue(sem, rgb)
    assert np.array_equal(lab_u8.cpu().numpy(), fx["label_u8"])  # incl. the rows without mass -> class 0 -> label 1
    assert np.array_equal(rgb_u8.cpu().numpy(), fx["rgb_u8"])
    _, bgr_u8 = ops.label_epilogue(sem, rgb, bgr=True, want_labels=False)
    assert np.array_equal(bgr_u8.cpu().numpy(), fx["rgb_u8"][:, ::-1])
    # a large, tie-heavy case against the torch expression of the reference
    g = torch.Generator().manual_seed(3)
    big = (torch.randint(0, 4, (5000, 40), generator=g).float() / 3).to(DEV)
    ref = big.clone()
    ref[ref.sum(-1) == 0] = 1
    ref = ref / ref.sum(-1, keepdim=True)
    is_max = ref == ref.max(-1, keepdim=True).values
    first_max = torch.where(is_max, torch.arange(40, device=DEV).expand_as(ref), torch.full_like(ref, 99).long()).min(-1).values  # lowest index among equal maxima
    lab, _ = ops.label_epilogue(big)
    assert torch.equal(lab.long(), first_max + 1)


@pytest.mark.gpu
def test_view_sampler_feeds_the_engine_batch_format(fx):
    from ucsa_neural_rendering_b200.frontend import ViewSampler

    h, w = int(fx["height"]), int(fx["width"])
    vs = ViewSampler(fx["ngp_poses"], fx["intrinsics"], h, w, images_h=fx["image_h"][None].repeat(6, 0),
                     labels=fx["labels"][None].repeat(6, 0), depths=fx["depth"][None].repeat(6, 0))
    g = torch.Generator(device=DEV).manual_seed(0)
    b = vs.sample(1, 500, generator=g)
    inds = b["inds"].cpu().numpy()
    assert b["rays_d"].shape == (500, 3) and b["gt_rgb"].dtype == torch.float16 and b["labels"].dtype == torch.int64
    np.testing.assert_allclose(b["rays_d"].cpu().numpy(), fx["rays_d"][1][inds], rtol=0, atol=2.5e-7)
    assert np.array_equal(b["gt_rgb"].cpu().numpy(), fx["image_h"].reshape(3, -1).T[inds])
    assert np.array_equal(b["depth"].cpu().numpy(), fx["depth"].reshape(-1)[inds])
    o, d, n = vs.full_view(0)
    assert d.shape == (h * w, 3) and torch.equal(o[0], vs.poses[0, :3, 3])


@pytest.mark.gpu
def test_pseudo_label_writer_async_png(golden_dir, tmp_path):
    """row f3, second half: label map / colour image / colour visualisation leave the device as u8, are copied
    asynchronously and PNG-encoded by worker threads; what lands on disk is what the reference's predict step writes
    (joint_train_lightning_net.py:755-782)."""
    import cv2

    from oracle import frontend as oracle_frontend
    from ucsa_neural_rendering_b200.labels import PseudoLabelWriter, default_palette

    h, w, c = 48, 64, 40
    g = torch.Generator().manual_seed(3)
    views = []
    with PseudoLabelWriter(str(tmp_path), h, w, workers=3, slots=2) as wr:  # fewer slots than views: back-pressure
        for v in range(5):
            sem = torch.rand(h * w, c, generator=g) ** 3
            sem[::11] = 0  # pixels without semantic mass -> uniform -> label 1
            rgb = torch.rand(h * w, 3, generator=g)
            views.append((sem, rgb))
            wr.submit(sem.cuda(), rgb.cuda(), f"{v:04d}", subfolder="novel_viewpoints" if v == 4 else "")
    assert wr.submitted == 5
    pal = default_palette()
    for v, (sem, rgb) in enumerate(views):
        sub = "novel_viewpoints" if v == 4 else ""
        label_ref, rgb_ref = oracle_frontend.label_epilogue(sem.numpy(), rgb.numpy())
        label = cv2.imread(str(tmp_path / sub / "nerf_label" / f"{v:04d}.png"), cv2.IMREAD_UNCHANGED)
        assert label.dtype == np.uint8 and label.shape == (h, w)
        assert np.array_equal(label, label_ref.reshape(h, w))
        img = cv2.imread(str(tmp_path / sub / "nerf_image" / f"{v:04d}.png"), cv2.IMREAD_COLOR)
        assert np.array_equal(img[..., ::-1], rgb_ref.reshape(h, w, 3))  # cv2 reads BGR back
        vis = cv2.imread(str(tmp_path / sub / "nerf_label_vis" / f"{v:04d}.png"), cv2.IMREAD_COLOR)
        assert np.array_equal(vis[..., ::-1], pal[label_ref.reshape(h, w)])
