"""CPU: the oracle's voxelisation / coordinate-map / kernel-map conventions.
MinkowskiEngine is absent, so these are property tests + hand-computed cases (parity unpinned)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import voxel as ov
from tests.helpers import random_surface_cloud, random_voxels


def test_quantize_is_float32_floor_division():
    pts = np.array([[0.049999, -0.05, 0.1], [-0.000001, 0.15, -0.1000001]], np.float32)
    q = ov.quantize_coords(pts, 0.05)
    ref = np.floor(pts / np.float32(0.05)).astype(np.int32)
    assert np.array_equal(q, ref)
    assert q[0, 1] == -1 and q[1, 0] == -1  # negatives floor toward -inf
    # float32 vs float64 division disagree on some inputs: the oracle must be the float32 one
    c = np.arange(-1200, 1200)
    p = (c * 0.05).astype(np.float32)
    assert (np.floor(p / np.float32(0.05)) != np.floor(p.astype(np.float64) / 0.05)).any()


def test_sparse_quantize_known_answer():
    pts = np.array([[0.01, 0.01, 0.01], [0.12, 0.0, 0.0], [0.02, 0.03, 0.04], [0.11, 0.01, 0.0], [-0.01, 0, 0]], np.float32)
    labels = np.array([3, 5, 3, 6, 2], np.int32)
    feats = np.arange(5, dtype=np.float32).reshape(5, 1)
    q, f, cl, um, inv = ov.sparse_quantize(pts, feats, labels, -1, True, True, False, 0.05)
    assert q.tolist() == [[0, 0, 0], [2, 0, 0], [-1, 0, 0]]
    assert um.tolist() == [0, 1, 4] and inv.tolist() == [0, 1, 0, 1, 2]
    assert f.reshape(-1).tolist() == [0, 1, 4]
    assert cl.tolist() == [3, -1, 2]  # voxel 1 holds labels 5 and 6 -> ignore


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 400), st.integers(0, 2 ** 31 - 1))
def test_unique_inverse_properties(n, seed):
    rng = np.random.default_rng(seed)
    c = rng.integers(-5, 5, (n, 3)).astype(np.int32)
    um, inv = ov.unique_first_occurrence(c)
    assert np.all(np.diff(um) > 0)
    assert np.array_equal(c[um][inv], c)
    assert len(np.unique(c, axis=0)) == len(um)
    # first occurrence: no earlier row holds the same voxel
    for u, r in enumerate(um):
        assert not (c[:r] == c[r]).all(1).any()


def test_stride_map_and_k2_pairs():
    rng = np.random.default_rng(0)
    c = random_voxels(rng, 3000)
    c2, parent = ov.stride_coords(c, 2)
    assert np.array_equal(c2[parent][:, 0], c[:, 0])
    assert np.array_equal(c2[parent][:, 1:], np.floor_divide(c[:, 1:], 2) * 2)
    maps = ov.kernel_map(c, c2, 2, 1)
    assert sum(len(i) for i, _ in maps) == len(c)  # each fine voxel in exactly one pair
    offs = ov.kernel_offsets(2, 1)
    for k, (i, o) in enumerate(maps):
        assert np.all(c[i, 1:] - c2[o, 1:] == offs[k])
        assert np.array_equal(parent[i], o)
    tmaps = ov.transposed_kernel_map(c, c2, 2, 1)
    for (i, o), (ti, to) in zip(maps, tmaps):
        assert sorted(zip(i.tolist(), o.tolist())) == sorted(zip(to.tolist(), ti.tolist()))
        assert np.all(np.diff(to) > 0)


def test_kernel_offsets_order():
    o3 = ov.kernel_offsets(3, 2)
    assert o3[0].tolist() == [-2, -2, -2] and o3[1].tolist() == [0, -2, -2] and o3[13].tolist() == [0, 0, 0]
    assert np.array_equal(o3[::-1], -o3)  # index reversal mirrors the offset (used by dgrad)
    o2 = ov.kernel_offsets(2, 4)
    assert o2.tolist() == [[0, 0, 0], [4, 0, 0], [0, 4, 0], [4, 4, 0], [0, 0, 4], [4, 0, 4], [0, 4, 4], [4, 4, 4]]


def test_k3_map_symmetry_and_centre():
    rng = np.random.default_rng(1)
    c = random_voxels(rng, 2000, span=10)
    maps = ov.kernel_map(c, c, 3, 1)
    assert np.array_equal(maps[13][0], np.arange(len(c))) and np.array_equal(maps[13][1], np.arange(len(c)))
    for k in range(13):
        a = set(zip(maps[k][0].tolist(), maps[k][1].tolist()))
        b = set(zip(maps[26 - k][1].tolist(), maps[26 - k][0].tolist()))
        assert a == b


def test_collation_batch_column_first():
    a = np.array([[1, 2, 3]], np.int32)
    b = np.array([[4, 5, 6], [7, 8, 9]], np.int32)
    assert ov.batched_coordinates([a, b]).tolist() == [[0, 1, 2, 3], [1, 4, 5, 6], [1, 7, 8, 9]]


def test_synthetic_scan_shapes():
    from lidog_b200.lidog import synth
    pts, lab = synth.make_scan(1234, "kitti")
    assert 100_000 < len(pts) < 140_000 and pts.dtype == np.float32 and lab.min() >= -1 and lab.max() <= 6
    pts2, _ = synth.make_scan(1234, "kitti")
    assert np.array_equal(pts, pts2)  # deterministic
    q = ov.quantize_coords(pts, 0.05)
    assert np.abs(q[:, :2]).max() < 1200 and q[:, 2].min() >= -200 and q[:, 2].max() < 160
    pts, _ = synth.make_scan(7, "nuscenes")
    assert 25_000 < len(pts) < 40_000


def test_bev_label_image_matches_numpy_restatement():
    """`bev_label_image` (the batched torch version the trainer runs, CPU here / CUDA in tests/test_gpu_trainer.py)
    against the numpy restatement of PC2ImgConverter.getBEVImageNew (semantickitti_bev.py:433-464): float32
    subtract, TRUE float32 division by the grid size, floor, last writer wins."""
    import torch
    from lidog_b200.lidog import synth
    from lidog_b200.lidog.step import bev_label_image
    for seed, shape in ((3, "nuscenes"), (4, "kitti")):
        cfg = synth.SHAPES[shape]
        pts, lab = synth.make_scan(seed, shape)
        q, _, colab, _, _ = ov.sparse_quantize(pts, np.ones((len(pts), 1), np.float32), lab, -1, True, True, False, 0.05)
        want = synth.make_bev_labels(q, colab, cfg["bound"], cfg["bev_img"])
        c4 = torch.from_numpy(np.concatenate([np.zeros((len(q), 1), np.int32), q], 1))
        got = bev_label_image(c4, torch.from_numpy(colab), 1, cfg["bound"], cfg["bev_img"])
        assert got.shape == (1, cfg["bev_img"], cfg["bev_img"])
        assert np.array_equal(got[0].numpy(), want)


def test_mix3d_shaped_sample_requantises_through_float32():
    """Mix3D-shaped synthetic samples (lidog_b200/lidog/synth.py, reference utils/datasets/mix3D.py:43-87): the
    merged cloud is metric float32 voxel corners of two scans; re-quantising it does NOT give back the union of
    the two integer voxel sets (float32(c * 0.05) / 0.05 floors to c - 1 for some c), and the oracle reproduces
    numpy's float32 arithmetic exactly."""
    from lidog_b200.lidog import synth
    pts, lab = synth.make_scan(21, "mix3d")
    assert pts.dtype == np.float32 and lab.dtype == np.int32 and len(pts) == len(lab)
    q, _, _, umap, inv = ov.sparse_quantize(pts, np.ones((len(pts), 1), np.float32), lab, -1, True, True, False, 0.05)
    want = np.floor(pts / np.float32(0.05)).astype(np.int32)
    assert np.array_equal(q[inv], want) and np.array_equal(want[umap], q)
    # the two sources, quantised on their own
    srcs = [synth._quantize_first(*synth.make_scan(21 + 5003 * j, "kitti"), 0.05)[0] for j in range(2)]
    union = np.unique(np.concatenate(srcs), axis=0)
    merged = np.unique(q, axis=0)
    assert len(merged) != len(union) or not np.array_equal(merged, union)  # the requantisation trap is real
    again, _ = synth.make_scan(21, "mix3d")
    assert np.array_equal(again, pts)  # deterministic


def test_quantize_keeps_float64_clouds_in_float64():
    """numpy divides `float64_array / python_float` in float64 (the reference's augmented training clouds,
    utils/common/augmentation.py:10-20) and `float32_array / python_float` in float32; the two floor differently on
    cell boundaries, and the oracle follows the array's precision."""
    rng = np.random.default_rng(3)
    c = rng.integers(-1200, 1200, 30000)
    p64 = np.stack([c * 0.05, c * 0.05 + 1e-9, c * 0.05 - 1e-9], 1)  # points on and next to cell boundaries
    q64 = ov.quantize_coords(p64, 0.05)
    assert np.array_equal(q64, np.floor(p64 / 0.05).astype(np.int32))
    q32 = ov.quantize_coords(p64.astype(np.float32), 0.05)
    assert np.array_equal(q32, np.floor(p64.astype(np.float32) / np.float32(0.05)).astype(np.int32))
    assert (q64 != q32).any()
    per_axis = ov.quantize_coords(p64, [0.3, 0.2, 1.0])
    assert np.array_equal(per_axis, np.floor(p64 / np.array([0.3, 0.2, 1.0])).astype(np.int32))


@pytest.mark.parametrize("n", [10_000, 200_000])
def test_plane_cloud_keeps_lidar_like_neighbourhoods(n):
    """The micro-benchmark clouds (BASELINE configs[3]) keep 5-12 occupied neighbours per voxel at any size."""
    from lidog_b200.lidog import synth
    pts, lab = synth.make_plane_cloud(n, 3)
    assert pts.dtype == np.float32 and lab.dtype == np.int32 and abs(len(pts) - n) < 12
    q = np.unique(ov.quantize_coords(pts, 0.05), axis=0)
    c4 = np.concatenate([np.zeros((len(q), 1), np.int32), q], 1)
    pairs = sum(len(i) for i, _ in ov.kernel_map(c4, c4, 3, 1))
    per_voxel = pairs / len(q) - 1  # without the centre
    assert 4.0 < per_voxel < 13.0, per_voxel
