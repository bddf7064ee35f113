"""Device side of the per-scan data path that precedes voxelisation: radius crop, random
subsampling, augmentation, bounds filter.  The reference runs these in numpy inside the DataLoader workers
(utils/datasets/semantickitti_bev.py:187,209-224, utils/datasets/dataset.py:58-72,
utils/common/augmentation.py:7-55); here they are torch ops on whatever device the cloud lives on, so a scan that
is already in HBM goes crop -> augment -> filter -> `sparse_quantize_batch` without a round trip to the host.

Arithmetic follows numpy exactly where a voxel index depends on it:
  * the crop compares `x*x + y*y + z*z` (float32, left to right, as np.sum over 3 squared columns) with in_R**2;
  * the augmentation is `points @ R` in FLOAT64 followed by per-axis float64 scales -- the reference multiplies the
    float32 cloud by scipy's float64 rotation, so everything downstream (bounds test, voxelisation) sees float64;
  * the random draws themselves (rotation axis / angle, scales, the subsampling permutation) stay on the host: they
    are a handful of numbers, and keeping numpy's generator keeps runs comparable with the reference.
"""
from __future__ import annotations

import numpy as np
import torch

GRID_BOUNDS = ((-60.0, 60.0), (-60.0, 60.0), (-10.0, 8.0))  # semantickitti_bev.py:137
EGO_BOX = ((-3.0, 3.0), (-2.0, 2.0))                        # semantickitti_bev.py:166-168


def radius_mask(points: torch.Tensor, in_radius: float) -> torch.Tensor:
    """`np.sum(np.square(points), axis=1) < in_R ** 2` (semantickitti_bev.py:187)."""
    p = points[:, :3]
    d2 = p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2]
    return d2 < in_radius ** 2


def bounds_mask(points: torch.Tensor, grid_bounds=GRID_BOUNDS, ego_box=EGO_BOX) -> torch.Tensor:
    """`filter_bounds` (semantickitti_bev.py:155-172): strictly inside the grid, outside the ego box."""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    (x0, x1), (y0, y1), (z0, z1) = grid_bounds
    inside = (x0 < x) & (x < x1) & (y0 < y) & (y < y1) & (z0 < z) & (z < z1)
    (ex0, ex1), (ey0, ey1) = ego_box
    ego = (ex0 < x) & (x < ex1) & (ey0 < y) & (y < ey1)
    return inside & ~ego


def rotation_matrix(axis, theta) -> np.ndarray:
    """expm(cross(eye(3), axis / |axis| * theta)) (augmentation.py:9-10) in closed form (Rodrigues): the matrix
    exponential of a skew-symmetric matrix.  float64, host side."""
    axis = np.asarray(axis, np.float64).reshape(3)
    theta = float(np.asarray(theta).reshape(-1)[0])
    a = axis / np.linalg.norm(axis) * theta
    K = np.cross(np.eye(3), a)  # the same generator matrix the reference exponentiates
    t = np.linalg.norm(a)
    if t < 1e-300:
        return np.eye(3)
    return np.eye(3) + (np.sin(t) / t) * K + ((1.0 - np.cos(t)) / (t * t)) * (K @ K)


def draw_augmentation(rng=np.random, scale_min=0.95, scale_max=1.05):
    """The random numbers of RandomRotation + RandomScale, drawn in the reference's order
    (augmentation.py:14-16,32-35): axis = rand(3) - 0.5, theta = pi/4 * (rand(1) - 0.5), then s_x, s_y, s_z."""
    axis = rng.rand(3) - 0.5
    theta = np.pi / 4 * (rng.rand(1) - 0.5)
    R = rotation_matrix(axis, theta)
    s = [(scale_max - scale_min) * rng.rand(1) + scale_min for _ in range(3)]
    return R, np.concatenate(s)


def augment(points: torch.Tensor, R=None, scale=None) -> torch.Tensor:
    """`coords @ R` then `coords[:, i] *= s_i` (augmentation.py:17-20,37-39), in float64 like numpy's
    float32 @ float64 promotion.  Returns a float64 tensor on the cloud's device."""
    p = points[:, :3].to(torch.float64)
    if R is not None:
        p = p @ torch.as_tensor(np.asarray(R, np.float64), device=p.device)
    if scale is not None:
        p = p * torch.as_tensor(np.asarray(scale, np.float64).reshape(1, 3), device=p.device)
    return p


def subsample_indices(num_points: int, sub_p, rng=np.random) -> np.ndarray:
    """`random_sample` (dataset.py:58-72): int(sub_p * N) indices without replacement, host generator."""
    if sub_p is None:
        return np.arange(num_points)
    return rng.choice(np.arange(num_points), int(sub_p * num_points), replace=False)


def prepare_scan(points: torch.Tensor, labels: torch.Tensor, in_radius: float = 50.0, R=None, scale=None,
                 sampled_idx=None, use_bounds: bool = True):
    """One scan, in the order of `__getitem__` (semantickitti_bev.py:186-224): radius crop, [subsample, augment],
    bounds filter.  -> (points, labels, kept) with `kept` = indices into the input cloud (the reference's
    `sampled_idx` after both filters); points are float64 when an augmentation was applied, else unchanged."""
    dev = points.device
    kept = torch.nonzero(radius_mask(points, in_radius)).squeeze(1)
    pts, lab = points[kept, :3], labels[kept]
    if sampled_idx is not None:
        idx = torch.as_tensor(np.asarray(sampled_idx), device=dev, dtype=torch.long)
        pts, lab, kept = pts[idx], lab[idx], kept[idx]
    if R is not None or scale is not None:
        pts = augment(pts, R, scale)
    if use_bounds:
        m = bounds_mask(pts)
        pts, lab, kept = pts[m], lab[m], kept[m]
    return pts, lab, kept


def mix3d_merge(source0: dict, source1: dict, voxel_size: float, ignore_label, ME=None) -> dict:
    """Mix3D merge of two already voxelised samples (`merge_data`, utils/datasets/mix3D.py:43-87) on the samples'
    device: the integer voxel coordinates of both samples go back to metric float32 (`coordinates * voxel_size`, a
    torch int tensor times a python float = a float32 product, `:45-46`), are concatenated (`:60`) and re-quantised
    with one hash build (`:67-72`); features / labels / sampled indices follow the first point of every merged voxel
    (`:74-76`).  The float32 round trip makes floor(float32(c * 0.05) / 0.05) = c - 1 for ~6 % of the integers
    (SURVEY.md 8a), so this is NOT the set union of the two voxel sets -- the arithmetic is reproduced, not shortcut.
    `source*` = dict(coordinates int [N,3], features [N,C], sem_labels [N], sampled_idx [N], xyz, idx)."""
    if ME is None:
        from lidog_b200 import me as ME
    c0, c1 = source0["coordinates"], source1["coordinates"]
    metric = torch.cat([c0.to(torch.float32) * voxel_size, c1.to(torch.float32) * voxel_size], dim=0)
    features = torch.cat([source0["features"], source1["features"]], dim=0)
    sem_labels = torch.cat([source0["sem_labels"], source1["sem_labels"]], dim=0)
    sampled_idx = torch.cat([source0["sampled_idx"], source1["sampled_idx"]], dim=0)
    q, _, _, voxel_idx = ME.utils.sparse_quantize(metric, features, labels=sem_labels, ignore_label=ignore_label,
                                                  quantization_size=voxel_size, return_index=True)
    voxel_idx = torch.as_tensor(voxel_idx, device=features.device)
    return dict(coordinates=torch.as_tensor(q), xyz=torch.cat([source0["xyz"], source1["xyz"]], dim=0),
                features=features[voxel_idx], sem_labels=sem_labels[voxel_idx],
                idx=torch.cat([source0["idx"].view(1, -1), source1["idx"].view(1, -1)], dim=0),
                sampled_idx=sampled_idx[voxel_idx])
