"""Synthetic SemanticKITTI- / nuScenes-shaped LiDAR scans (numpy, deterministic).

The benchmark inputs SURVEY.md section 8(d) specifies: a spinning LiDAR
ray-cast onto a ground plane, two street-canyon walls and a set of boxes, with
range noise, followed by LiDOG's bounds filter (reference
utils/datasets/semantickitti_bev.py:155-172: |x|,|y| < 60 m, -10 < z < 8 m,
ego box removed) and the 50 m radius crop (`:187`).  Labels are uniform per
surface in {-1, 0 .. C-1}.  No file IO; this replaces the dataset readers, which
are out of scope.
"""
from __future__ import annotations

import numpy as np

SENSOR_HEIGHT = 1.73

SHAPES = {
    # beams, azimuth steps, elevation range (deg), BEV bound (m), BEV label image size
    "kitti": dict(beams=64, azimuth=2083, el=(-24.8, 2.0), bound=50.0, bev_img=167),
    "nuscenes": dict(beams=32, azimuth=1090, el=(-30.0, 10.0), bound=30.0, bev_img=100),
    # BASELINE configs[4]: two kitti-shaped scans merged and re-quantised the way utils/datasets/mix3D.py:43-87 does
    "mix3d": dict(beams=64, azimuth=2083, el=(-24.8, 2.0), bound=50.0, bev_img=167, mix_of="kitti"),
}


def _ray_dirs(beams, azimuth, el_range):
    el = np.deg2rad(np.linspace(el_range[0], el_range[1], beams, dtype=np.float64))
    az = np.linspace(-np.pi, np.pi, azimuth, endpoint=False, dtype=np.float64)
    el, az = np.meshgrid(el, az, indexing="ij")
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=-1)
    return d.reshape(-1, 3)


def _quantize_first(points: np.ndarray, labels: np.ndarray, voxel_size: float):
    """The dataset-side sparse_quantize of ONE source scan, reduced to what the merge needs: integer voxel
    coordinates (first occurrence order) and the first point's label (semantickitti_bev.py:232-244)."""
    q = np.floor(points / np.float32(voxel_size)).astype(np.int32)
    _, first = np.unique(q, axis=0, return_index=True)
    first = np.sort(first)
    return q[first], labels[first]


def make_mix3d_scan(seed: int, num_classes: int = 7, voxel_size: float = 0.05, source: str = "kitti"):
    """Mix3D-shaped sample (utils/datasets/mix3D.py:43-87): two source scans, each already voxelised by its
    dataset, are turned back into metric float32 coordinates (`coordinates * voxel_size`, `:45-46`), concatenated
    (`:60`) and handed to sparse_quantize again (`:67-72`).  float32(c * 0.05) / 0.05 floors to c - 1 for ~6 %
    of the integers (SURVEY.md section 8a note), so the merged voxel set is NOT the union of the two source sets;
    returning the metric points keeps that arithmetic inside the (GPU) voxelisation under test."""
    parts, labs = [], []
    for j in range(2):
        pts, lab = make_scan(seed + 5003 * j, source, num_classes)
        q, l = _quantize_first(pts, lab, voxel_size)
        parts.append((q.astype(np.float32) * np.float32(voxel_size)).astype(np.float32))
        labs.append(l)
    return np.concatenate(parts, 0), np.concatenate(labs, 0).astype(np.int32)


def make_scan(seed: int, shape: str = "kitti", num_classes: int = 7, max_range: float = 50.0):
    """-> (points float32 [N,3], labels int32 [N]) after LiDOG's crop/filters."""
    cfg = SHAPES[shape]
    if "mix_of" in cfg:
        return make_mix3d_scan(seed, num_classes, source=cfg["mix_of"])
    rng = np.random.default_rng(seed)
    d = _ray_dirs(cfg["beams"], cfg["azimuth"], cfg["el"])
    n = d.shape[0]
    t_best = np.full(n, np.inf)
    surf = np.full(n, -1, np.int64)
    eps = 1e-9

    # surface 0: ground plane
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(d[:, 2] < -eps, -SENSOR_HEIGHT / d[:, 2], np.inf)
    t_best, surf = t, np.where(np.isfinite(t), 0, surf)

    # surfaces 1,2: canyon walls y = +-wy, finite height
    sid = 1
    for sign in (+1.0, -1.0):
        wy = sign * rng.uniform(7.0, 14.0)
        height = rng.uniform(4.0, 9.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(d[:, 1] * sign > eps, wy / d[:, 1], np.inf)
        z = t * d[:, 2]
        ok = (z > -SENSOR_HEIGHT) & (z < height - SENSOR_HEIGHT)
        # leave gaps (side streets) in the walls
        x = t * d[:, 0]
        gap_c = rng.uniform(-40, 40, size=3)
        for g in gap_c:
            ok &= np.abs(x - g) > 4.0
        t = np.where(ok, t, np.inf)
        better = t < t_best
        t_best = np.where(better, t, t_best)
        surf = np.where(better, sid, surf)
        sid += 1

    # boxes (cars, poles, kiosks) standing on the ground
    n_boxes = 48
    for _ in range(n_boxes):
        cx, cy = rng.uniform(-45, 45), rng.uniform(-6.5, 6.5)
        if abs(cx) < 5 and abs(cy) < 3:
            cx += 8.0
        sx, sy, sz = rng.uniform(0.3, 4.5), rng.uniform(0.3, 2.0), rng.uniform(0.8, 3.0)
        lo = np.array([cx - sx / 2, cy - sy / 2, -SENSOR_HEIGHT])
        hi = np.array([cx + sx / 2, cy + sy / 2, -SENSOR_HEIGHT + sz])
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            t0 = lo * inv
            t1 = hi * inv
        tmin = np.nanmax(np.minimum(t0, t1), axis=1)
        tmax = np.nanmin(np.maximum(t0, t1), axis=1)
        hit = (tmax >= tmin) & (tmin > 0.5)
        t = np.where(hit, tmin, np.inf)
        better = t < t_best
        t_best = np.where(better, t, t_best)
        surf = np.where(better, sid, surf)
        sid += 1

    keep = np.isfinite(t_best)
    t_best = np.where(keep, t_best, 1.0) + rng.normal(0.0, 0.02, size=n)
    pts = d * t_best[:, None]
    keep &= (t_best > 0.5) & (np.sum(pts * pts, axis=1) < max_range ** 2)

    # LiDOG bounds filter + ego box
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    keep &= (-60 < x) & (x < 60) & (-60 < y) & (y < 60) & (-10 < z) & (z < 8)
    keep &= ~((-3 < x) & (x < 3) & (-2 < y) & (y < 2))

    surf_label = rng.integers(-1, num_classes, size=sid).astype(np.int32)
    labels = surf_label[np.maximum(surf, 0)]
    return pts[keep].astype(np.float32), labels[keep].astype(np.int32)


def make_bev_labels(qcoords: np.ndarray, colabels: np.ndarray, bound: float, img: int,
                    voxel_size: float = 0.05) -> np.ndarray:
    """BEV label image the way LiDOG's dataset builds it
    (PC2ImgConverter.getBEVImageNew, semantickitti_bev.py:433-464): int64 [img, img],
    -1 = ignore; last writer wins."""
    pts = (qcoords * voxel_size).astype(np.float32)
    grid = (bound - (-bound)) / img
    lab = -np.ones((img, img), np.int64)
    valid = colabels != -1
    p, l = pts[valid], colabels[valid]
    inb = (-bound < p[:, 0]) & (p[:, 0] < bound) & (-bound < p[:, 1]) & (p[:, 1] < bound) & (-10 < p[:, 2]) & (p[:, 2] < 8)
    p, l = p[inb], l[inb]
    px = np.floor((p[:, 0] - (-bound)) / grid).astype(np.int64)
    py = np.floor(img - (p[:, 1] - (-bound)) / grid).astype(np.int64) - 1
    lab[py, px] = l
    return lab


def make_batch(batch_size: int, seed: int = 1234, shape: str = "kitti", num_classes: int = 7):
    """List of `batch_size` scans, a different scan per batch slot."""
    return [make_scan(seed + i, shape, num_classes) for i in range(batch_size)]


def make_plane_cloud(n_points: int, seed: int = 0, extent: float = 40.0, planes: int = 12, noise: float = 0.01,
                     num_classes: int = 7):
    """Micro-benchmark cloud (BASELINE configs[3]): `n_points` points on a few random planes,
    which keeps the 5-12 occupied neighbours per 0.05 m voxel that LiDAR surfaces have at any point count from
    10 k to 2 M.  -> (points float32 [N,3], labels int32 [N]); the plane sizes grow with N so the point density per
    voxel stays ~1-2."""
    rng = np.random.default_rng(seed)
    per = max(n_points // planes, 1)
    # a plane of side L holds (L / 0.05)^2 voxels; aim at ~1.5 points per voxel
    side = float(np.clip(0.05 * np.sqrt(per / 1.5), 1.0, extent))
    pts, labs = [], []
    for j in range(planes):
        origin = rng.uniform(-extent / 2, extent / 2, 3)
        u, v = rng.standard_normal(3), rng.standard_normal(3)
        u /= np.linalg.norm(u)
        v -= u * (u @ v)
        v /= np.linalg.norm(v)
        a, b = rng.uniform(-side / 2, side / 2, (2, per))
        pts.append(origin + a[:, None] * u + b[:, None] * v + rng.normal(0.0, noise, (per, 3)))
        labs.append(np.full(per, rng.integers(-1, num_classes), np.int32))
    return np.concatenate(pts).astype(np.float32), np.concatenate(labs)
