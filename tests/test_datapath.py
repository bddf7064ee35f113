"""Device data path helpers (lidog_b200/lidog/datapath.py) against numpy restatements of the
reference's dataset code, on CPU tensors (the same torch ops run on the GPU).  Masks and indices bit exact, the
float64 augmentation to the last ulp of a 3-term dot product."""
import numpy as np
import torch

from lidog_b200.lidog import datapath as dp
from lidog_b200.lidog import synth


def _np_filter_bounds(points):  # semantickitti_bev.py:155-172
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    inb = np.logical_and(np.logical_and(-60 < x, x < 60), np.logical_and(np.logical_and(-60 < y, y < 60),
                                                                          np.logical_and(-10 < z, z < 8)))
    ego = np.logical_not(np.logical_and(np.logical_and(-3 < x, x < 3), np.logical_and(-2 < y, y < 2)))
    return np.logical_and(inb, ego)


def _raw_cloud(seed, n=50000):
    rng = np.random.default_rng(seed)
    pts = (rng.standard_normal((n, 3)) * np.array([30.0, 30.0, 4.0])).astype(np.float32)
    pts[:2000] = (rng.standard_normal((2000, 3)) * np.array([2.0, 1.5, 1.0])).astype(np.float32)  # around the ego box
    return pts, rng.integers(-1, 7, n).astype(np.int32)


def test_radius_and_bounds_masks_match_numpy():
    pts, _ = _raw_cloud(0)
    t = torch.from_numpy(pts)
    assert np.array_equal(dp.radius_mask(t, 50.0).numpy(), np.sum(np.square(pts), axis=1) < 50.0 ** 2)
    assert np.array_equal(dp.bounds_mask(t).numpy(), _np_filter_bounds(pts))
    p64 = pts.astype(np.float64) * 1.0000001
    assert np.array_equal(dp.bounds_mask(torch.from_numpy(p64)).numpy(), _np_filter_bounds(p64))


def test_rotation_matrix_is_the_matrix_exponential():
    from scipy.linalg import expm, norm
    rng = np.random.RandomState(3)
    for _ in range(20):
        axis, theta = rng.rand(3) - 0.5, np.pi / 4 * (rng.rand(1) - 0.5)
        want = expm(np.cross(np.eye(3), axis / norm(axis) * theta))  # augmentation.py:9-10
        got = dp.rotation_matrix(axis, theta)
        assert np.allclose(got, want, rtol=0, atol=1e-15)
    assert np.array_equal(dp.rotation_matrix([1.0, 0.0, 0.0], [0.0]), np.eye(3))


def test_augment_is_float64_like_numpy():
    pts, _ = _raw_cloud(1, 20000)
    R, s = dp.draw_augmentation(np.random.RandomState(5))
    want = pts @ R  # float32 @ float64 -> float64
    assert want.dtype == np.float64
    want[:, 0] *= s[0]
    want[:, 1] *= s[1]
    want[:, 2] *= s[2]
    got = dp.augment(torch.from_numpy(pts), R, s)
    assert got.dtype == torch.float64
    assert np.allclose(got.numpy(), want, rtol=4e-16, atol=0)  # (BLAS and torch may order the 3-term sums differently)
    same_draws = dp.draw_augmentation(np.random.RandomState(5))
    assert np.array_equal(same_draws[0], R) and np.array_equal(same_draws[1], s)


def test_prepare_scan_follows_getitem_order():
    pts, lab = _raw_cloud(2)
    rs = np.random.RandomState(7)
    # numpy restatement of __getitem__ (semantickitti_bev.py:186-224) with fixed draws
    mask = np.sum(np.square(pts), axis=1) < 50.0 ** 2
    p, l = pts[mask], lab[mask]
    idx = dp.subsample_indices(p.shape[0], 0.8, np.random.RandomState(11))
    assert len(idx) == int(0.8 * p.shape[0]) and len(np.unique(idx)) == len(idx)
    R, s = dp.draw_augmentation(rs)
    p2 = p[idx] @ R
    p2 = p2 * s.reshape(1, 3)
    l2 = l[idx]
    keep = _np_filter_bounds(p2)
    want_p, want_l = p2[keep], l2[keep]
    got_p, got_l, kept = dp.prepare_scan(torch.from_numpy(pts), torch.from_numpy(lab), 50.0, R, s, idx)
    assert got_p.dtype == torch.float64 and np.array_equal(got_l.numpy(), want_l)
    assert np.allclose(got_p.numpy(), want_p, rtol=4e-16, atol=0)
    assert np.array_equal(kept.numpy(), np.nonzero(mask)[0][idx][keep])
    # without augmentation the cloud keeps its dtype and only the two filters act
    q_p, q_l, _ = dp.prepare_scan(torch.from_numpy(pts), torch.from_numpy(lab), 50.0)
    m2 = mask.copy()
    m2[mask] = _np_filter_bounds(pts[mask])
    assert q_p.dtype == torch.float32 and np.array_equal(q_p.numpy(), pts[m2]) and np.array_equal(q_l.numpy(), lab[m2])


def test_synthetic_scans_already_satisfy_the_filters():
    pts, lab = synth.make_scan(9, "nuscenes")
    t = torch.from_numpy(pts)
    assert bool(dp.radius_mask(t, 50.0).all()) and bool(dp.bounds_mask(t).all())
