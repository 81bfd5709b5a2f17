"""Scene / pose side of the path (row f4 of SURVEY.md section 8): the on-disk format the reference trains from and the
pose conventions either side of render().  Host logic only (numpy); everything here is exercised on CPU.

* transforms_{train,test}.json as written by preprocessing_scripts/scannet2nerf.py:196-232 and read by
  nr4seg/dataset/scannet_ngp_joint.py:127-141: "h", "w", "fl_x", "fl_y", "cx", "cy", "one_m_to_scene_uom" and
  "frames" = [{"file_path", "label_path", "transform_matrix"}];
* nerf_matrix_to_ngp (nr4seg/dataset/ngp_utils.py:7-18): axis permutation + sign flips from the NeRF / Blender
  camera convention to instant-ngp's;
* novel view points between consecutive training poses (scannet_ngp_joint.py:229-262): the rotation half-way along
  the geodesic (Slerp at t = 0.5), the translation at the mid-point, closing the loop from the last to the first view.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np


def nerf_matrix_to_ngp(pose: np.ndarray) -> np.ndarray:
    """rows (1, 2, 0) of the 4x4 cam2world matrix, camera y and z axes negated; float32"""
    pose = np.asarray(pose)
    out = np.eye(4, dtype=np.float32)
    out[:3] = pose[[1, 2, 0]]
    out[:3, 1:3] *= -1.0
    return out


@dataclass
class SceneInfo:
    height: int
    width: int
    intrinsics: np.ndarray  # (fx, fy, cx, cy)
    one_m_to_scene_uom: float
    poses: np.ndarray  # [V,4,4] float32, instant-ngp convention
    image_paths: list
    label_paths: list
    depth_paths: list


def split_frames(frames, mode: str):
    """train: all but the last 20 %; val: the last 20 %; predict / anything else: all (scannet_ngp_joint.py:142-148)"""
    k = int(0.2 * len(frames))
    if mode == "val":
        return frames[-k:]
    if mode == "predict":
        return frames
    return frames[:-k]


def load_transforms(scene_root: str, mode: str = "train", file_name: str = "transforms_train.json") -> SceneInfo:
    with open(os.path.join(scene_root, file_name)) as fh:
        info = json.load(fh)
    frames = split_frames(info["frames"], mode)
    poses, images, labels, depths = [], [], [], []
    for f in frames:
        image_path = os.path.join(scene_root, f["file_path"])
        stem = os.path.basename(image_path).split(".")[0]
        images.append(image_path)
        labels.append(os.path.join(scene_root, f["label_path"]))
        depths.append(os.path.join(scene_root, "depth", stem + ".png"))
        poses.append(nerf_matrix_to_ngp(np.array(f["transform_matrix"], dtype=np.float32)))
    return SceneInfo(int(info["h"]), int(info["w"]),
                     np.array([info["fl_x"], info["fl_y"], info["cx"], info["cy"]]), float(info["one_m_to_scene_uom"]),
                     np.stack(poses, axis=0) if poses else np.zeros((0, 4, 4), np.float32), images, labels, depths)


def _rotation_log(r: np.ndarray) -> np.ndarray:
    """axis * angle of a rotation matrix (angle in [0, pi])"""
    cos = np.clip((np.trace(r) - 1.0) / 2.0, -1.0, 1.0)
    angle = np.arccos(cos)
    if angle < 1e-12:
        return np.zeros(3)
    if np.pi - angle < 1e-6:  # near pi the skew part vanishes: take the axis from the symmetric part
        b = (r + np.eye(3)) / 2.0
        axis = np.sqrt(np.clip(np.diag(b), 0.0, None))
        k = int(np.argmax(axis))
        axis = b[k] / axis[k]
        return axis / np.linalg.norm(axis) * angle
    w = np.array([r[2, 1] - r[1, 2], r[0, 2] - r[2, 0], r[1, 0] - r[0, 1]])
    return w / (2.0 * np.sin(angle)) * angle


def _rotation_exp(w: np.ndarray) -> np.ndarray:
    angle = np.linalg.norm(w)
    if angle < 1e-12:
        return np.eye(3)
    k = w / angle
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(angle) * kx + (1.0 - np.cos(angle)) * (kx @ kx)


def interpolate_novel_poses(poses) -> list:
    """One pose between every pair of consecutive poses and between the last and the first (NeRF convention in,
    NeRF convention out, float64 like the reference): rotation = Slerp at the half-way time, translation = mid-point."""
    poses = [np.asarray(p, dtype=np.float64) for p in poses]
    ring = poses + [poses[0]]
    out = []
    for a, b in zip(ring[:-1], ring[1:]):
        mid = np.eye(4)
        mid[:3, :3] = a[:3, :3] @ _rotation_exp(0.5 * _rotation_log(a[:3, :3].T @ b[:3, :3]))
        mid[:3, 3] = (a[:3, 3] + b[:3, 3]) / 2.0
        out.append(mid)
    return out
