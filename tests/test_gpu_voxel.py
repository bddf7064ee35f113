"""GPU parity (bit-exact): voxelisation, unique/inverse maps, stride maps, kernel maps vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import voxel as ov
from tests.helpers import random_surface_cloud, random_voxels

pytestmark = pytest.mark.gpu


def test_sparse_quantize_matches_oracle(cuda):
    import MinkowskiEngine as ME
    rng = np.random.default_rng(0)
    for n in (0 + 7, 1000, 60000):
        pts = random_surface_cloud(rng, n)
        labels = rng.integers(-1, 7, pts.shape[0]).astype(np.int32)
        feats = np.ones((pts.shape[0], 1), np.float32)
        ref = ov.sparse_quantize(pts, feats, labels, -1, True, True, False, 0.05)
        got = ME.utils.sparse_quantize(pts, feats, labels=labels, ignore_label=-1, quantization_size=0.05,
                                       return_index=True, return_inverse=True)
        assert len(got) == 5
        for g, r, name in zip(got, ref, ("coords", "feats", "colabels", "unique_map", "inverse_map")):
            assert np.array_equal(np.asarray(g), r), name
        q, um, inv = got[0], got[3], got[4]
        assert np.all(np.diff(um) > 0)                       # first-occurrence order
        assert np.array_equal(q[inv], ov.quantize_coords(pts, 0.05))  # q[unique][inverse] == q


def test_sparse_quantize_variants(cuda):
    import MinkowskiEngine as ME
    rng = np.random.default_rng(1)
    pts = random_surface_cloud(rng, 5000)
    # 4-tuple form used by mix3D.py:67-72 (no inverse), tensor input, per-axis size (minkunet_bev.py:279-284)
    labels = rng.integers(0, 3, pts.shape[0]).astype(np.int32)
    feats = rng.standard_normal((pts.shape[0], 2)).astype(np.float32)
    got = ME.utils.sparse_quantize(pts, feats, labels=labels, ignore_label=-1, quantization_size=0.05, return_index=True)
    ref = ov.sparse_quantize(pts, feats, labels, -1, True, False, False, 0.05)
    assert len(got) == 4 and all(np.array_equal(g, r) for g, r in zip(got, ref))
    got = ME.utils.sparse_quantize(torch.from_numpy(pts), quantization_size=[0.3, 0.2, 1.0], return_index=True,
                                   return_inverse=True)
    ref = ov.sparse_quantize(pts, quantization_size=[0.3, 0.2, 1.0], return_index=True, return_inverse=True)
    assert all(np.array_equal(g.numpy(), r) for g, r in zip(got, ref))
    um, inv = ME.utils.sparse_quantize(pts, quantization_size=0.05, return_maps_only=True, return_inverse=True)
    ref = ov.sparse_quantize(pts, quantization_size=0.05, return_maps_only=True, return_inverse=True)
    assert np.array_equal(um, ref[0]) and np.array_equal(inv, ref[1])


def _augmented_cloud(seed, n=40000):
    """A float64 cloud the way the reference's training path makes it: float32 points times a float64 rotation
    and per-axis scales (utils/common/augmentation.py:10-44)."""
    rng = np.random.default_rng(seed)
    pts = random_surface_cloud(rng, n)
    th = 0.3 * (rng.random() - 0.5)
    R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
    out = pts @ R  # float32 @ float64 -> float64
    out[:, 0] *= 0.95 + 0.1 * rng.random()
    out[:, 1] *= 0.95 + 0.1 * rng.random()
    assert out.dtype == np.float64
    # plus points exactly on / next to voxel boundaries, where the two precisions are known to floor differently
    c = rng.integers(-240, 240, (2000, 3))
    return np.concatenate([out, c * 0.05, c * 0.05 + 1e-9], 0)


def test_sparse_quantize_float64_clouds_divide_in_float64(cuda):
    """Augmented training clouds are float64 and numpy divides them in float64 (lg_quantize_points_f64); a float32
    division of the same points lands in a different voxel for some of them."""
    import MinkowskiEngine as ME
    pts = _augmented_cloud(5)
    labels = np.random.default_rng(6).integers(-1, 7, pts.shape[0]).astype(np.int32)
    ref = ov.sparse_quantize(pts, None, labels, -1, True, True, False, 0.05)
    got = ME.utils.sparse_quantize(pts, labels=labels, ignore_label=-1, quantization_size=0.05, return_index=True,
                                   return_inverse=True)
    for g, r, name in zip(got, ref, ("coords", "colabels", "unique_map", "inverse_map")):
        assert np.array_equal(np.asarray(g), r), name
    assert np.array_equal(got[0][got[3]], np.floor(pts / 0.05).astype(np.int32))
    c32, inv32 = ME.utils.sparse_quantize(pts.astype(np.float32), quantization_size=0.05, return_inverse=True)
    assert (c32[inv32] != got[0][got[3]]).any()  # per point: the precisions really differ on this cloud
    t = ME.utils.sparse_quantize(torch.from_numpy(pts).to(cuda), quantization_size=0.05)  # device float64 tensor
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), ref[0])


def test_mix3d_requantisation_trap(cuda):
    """floor(float32(c*0.05)/0.05) != c for some c (SURVEY 8a notes): GPU must follow the fp32 ops."""
    import MinkowskiEngine as ME
    c = np.arange(-1200, 1200, dtype=np.int32)
    pts = np.stack([c * 0.05, c * 0.05, np.zeros_like(c, dtype=np.float64)], 1).astype(np.float32)
    got = ME.utils.sparse_quantize(pts, quantization_size=0.05)
    ref = ov.sparse_quantize(pts, quantization_size=0.05)
    assert np.array_equal(got, ref)
    assert (ov.quantize_coords(pts, 0.05)[:, 0] != c).sum() > 0


def test_out_of_range_coordinate_is_an_error(cuda):
    import MinkowskiEngine as ME
    pts = np.array([[0, 0, 0], [40000 * 0.05, 0, 0]], np.float32)
    with pytest.raises(RuntimeError, match="range"):
        ME.utils.sparse_quantize(pts, quantization_size=0.05)


def _manager(coords, dev):
    from lidog_b200.me.coords import CoordinateManager
    return CoordinateManager(torch.from_numpy(coords).to(dev))


def test_stride_maps_match_oracle(cuda):
    rng = np.random.default_rng(2)
    coords = random_voxels(rng, 30000)
    cm = _manager(coords, cuda)
    assert not cm.had_duplicates
    assert np.array_equal(cm.get_coords(1).cpu().numpy(), coords)  # row order preserved
    cur = coords
    for ts in (2, 4, 8, 16):
        ref, inv = ov.stride_coords(cur, ts)
        lvl = cm.level(ts)
        assert np.array_equal(lvl.coords.cpu().numpy(), ref), ts
        assert np.array_equal(lvl.parent_of_finer.cpu().numpy(), inv), ts
        cur = ref


def test_duplicate_coordinates_keep_first(cuda):
    import MinkowskiEngine as ME
    rng = np.random.default_rng(3)
    coords = random_voxels(rng, 2000)
    dup = np.concatenate([coords, coords[::3]], 0)
    feats = torch.arange(dup.shape[0], dtype=torch.float32, device=cuda).view(-1, 1)
    st = ME.SparseTensor(coordinates=torch.from_numpy(dup).to(cuda), features=feats)
    assert st.F.shape[0] == coords.shape[0]
    assert np.array_equal(st.C.cpu().numpy(), coords)
    assert torch.equal(st.F.view(-1).cpu(), torch.arange(coords.shape[0], dtype=torch.float32))


@pytest.mark.parametrize("ksize,ts", [(3, 1), (5, 1), (3, 2), (3, 4)])
def test_kernel_map_pairs_match_oracle(cuda, ksize, ts):
    rng = np.random.default_rng(4)
    coords = random_voxels(rng, 20000)
    cm = _manager(coords, cuda)
    c = cm.get_coords(ts).cpu().numpy()
    plan = cm.plan("same", ts, ts, ksize)
    i, o, koff = cm.kernel_map_pairs(plan)
    i, o, koff = i.cpu().numpy(), o.cpu().numpy(), koff.numpy()
    ref = ov.kernel_map(c, c, ksize, ts)
    assert koff[-1] == sum(len(m[0]) for m in ref)
    for k, (ri, ro) in enumerate(ref):
        assert np.array_equal(i[koff[k]:koff[k + 1]], ri), k
        assert np.array_equal(o[koff[k]:koff[k + 1]], ro), k
    # every pair's coordinates differ by exactly off_k
    offs = ov.kernel_offsets(ksize, ts)
    for k in range(len(ref)):
        d = c[i[koff[k]:koff[k + 1]], 1:] - c[o[koff[k]:koff[k + 1]], 1:]
        assert np.all(d == offs[k])
    # tile masks agree with the table
    nbr = plan.nbr.cpu().numpy()
    mask = plan.tile_mask.cpu().numpy().view(np.uint32)
    for t in range(nbr.shape[1] // 128):
        for k in range(nbr.shape[0]):
            has = (nbr[k, t * 128:(t + 1) * 128] >= 0).any()
            assert bool((mask[t, k // 32] >> (k % 32)) & 1) == bool(has)


@pytest.mark.parametrize("kind,ksize,ts", [("same", 3, 1), ("same", 3, 2), ("down", 2, 1), ("down", 2, 2)])
def test_sorted_plan_is_a_row_permutation_of_the_natural_plan(cuda, kind, ksize, ts):
    """lg_kernel_map_sorted: identical pair set, out_row a permutation, tile masks consistent, deterministic."""
    rng = np.random.default_rng(6)
    coords = random_voxels(rng, 20000)
    cm = _manager(coords, cuda)
    ts_out = ts if kind == "same" else 2 * ts
    nat = cm.plan(kind, ts, ts_out, ksize)
    srt = cm.plan(kind + "_sorted", ts, ts_out, ksize)
    n_out = nat.n_out
    assert srt.n_out == n_out and srt.n_slots == nat.n_slots and srt.K == nat.K
    orow = srt.out_row.cpu().numpy()
    assert np.array_equal(np.sort(orow[:n_out]), np.arange(n_out)) and np.all(orow[n_out:] == -1)
    nn, sn = nat.nbr.cpu().numpy(), srt.nbr.cpu().numpy()
    assert np.array_equal(sn[:, :n_out], nn[:, orow[:n_out]])  # slot s gathers what row out_row[s] gathers
    assert np.all(sn[:, n_out:] == -1)
    mask = srt.tile_mask.cpu().numpy().view(np.uint32).reshape(-1)
    for t in range(sn.shape[1] // 128):
        for k in range(sn.shape[0]):
            assert bool((mask[t] >> k) & 1) == bool((sn[k, t * 128:(t + 1) * 128] >= 0).any())
    # fewer (tile, offset) units than the natural order, and the same plan when built again
    units = lambda m: int(sum(bin(int(v)).count("1") for v in m.reshape(-1)))
    assert units(mask) <= units(nat.tile_mask.cpu().numpy().view(np.uint32))
    again = cm._sorted_plan(cm.level(ts), cm.level(ts_out), ksize, ts)
    assert torch.equal(again.out_row, srt.out_row) and torch.equal(again.nbr, srt.nbr)


def test_stride2_plans_match_oracle(cuda):
    rng = np.random.default_rng(5)
    coords = random_voxels(rng, 20000)
    cm = _manager(coords, cuda)
    for ts in (1, 2):
        fine, coarse = cm.get_coords(ts).cpu().numpy(), cm.get_coords(2 * ts).cpu().numpy()
        down = cm.plan("down", ts, 2 * ts, 2)
        i, o, koff = (t.cpu().numpy() if hasattr(t, "cpu") else t for t in cm.kernel_map_pairs(down))
        ref = ov.kernel_map(fine, coarse, 2, ts)
        assert koff[-1] == fine.shape[0]  # every fine voxel appears in exactly one pair
        for k, (ri, ro) in enumerate(ref):
            assert np.array_equal(i[koff[k]:koff[k + 1]], ri) and np.array_equal(o[koff[k]:koff[k + 1]], ro)
        up = cm.plan("up", 2 * ts, ts, 2)
        g, orow = up.nbr.cpu().numpy(), up.out_row.cpu().numpy()
        mask = up.tile_mask.cpu().numpy().view(np.uint32).reshape(-1)
        reft = ov.transposed_kernel_map(fine, coarse, 2, ts)
        got = {k: [] for k in range(8)}
        for t in range(g.shape[0] // 128):
            sl = slice(t * 128, (t + 1) * 128)
            valid = orow[sl] >= 0
            if not valid.any():
                assert not (g[sl] >= 0).any()
                continue
            assert bin(int(mask[t])).count("1") == 1
            k = int(mask[t]).bit_length() - 1
            got[k].append(np.stack([g[sl][valid], orow[sl][valid]], 1))
        seen = 0
        for k, (ri, ro) in enumerate(reft):
            gk = np.concatenate(got[k]) if got[k] else np.zeros((0, 2), np.int64)
            assert np.array_equal(gk[:, 0], ri) and np.array_equal(gk[:, 1], ro), k
            seen += len(ro)
        assert seen == fine.shape[0]


def test_split_views_of_a_shared_index_equal_independent_managers(cuda):
    """Multi-source step (trainer_lighting_2d_multi.py:146-167): both source batches are voxelised and hashed ONCE;
    `CoordinateManager.split` hands each forward pass a view whose levels and gather plans must be bit-identical to
    those of a manager built from that source alone."""
    from lidog_b200.me.coords import CoordinateManager
    rng = np.random.default_rng(8)
    a, b = random_voxels(rng, 9000, batch=2), random_voxels(rng, 7000, batch=3)
    both = np.concatenate([a, b + np.array([[2, 0, 0, 0]], np.int32)], 0)
    shared = _manager(both, cuda)
    views = shared.split([2, 3])
    for view, src in zip(views, (a, b)):
        alone = _manager(src, cuda)
        for ts in (1, 2, 4, 8, 16):
            assert torch.equal(view.levels[ts].coords, alone.levels[ts].coords), ts
            if ts > 1:
                got = view.levels[ts].parent_of_finer - view.levels[ts].row_offset
                assert torch.equal(got, alone.levels[ts].parent_of_finer), ts
        for key in (("same", 1, 1, 3), ("same", 1, 1, 5), ("same_sorted", 1, 1, 3), ("same_sorted", 4, 4, 3),
                    ("down_sorted", 1, 2, 2), ("down", 2, 4, 2), ("up", 2, 1, 2), ("up", 8, 4, 2), ("identity", 2, 2, 1)):
            p, q = view.plan(*key), alone.plan(*key)
            assert (p.n_out, p.n_in, p.n_slots, p.K) == (q.n_out, q.n_in, q.n_slots, q.K), key
            assert torch.equal(p.nbr, q.nbr), key
            assert torch.equal(p.tile_mask, q.tile_mask), key
            assert (p.out_row is None) == (q.out_row is None) and (p.out_row is None or torch.equal(p.out_row, q.out_row)), key


def test_mix3d_merge_on_the_device_matches_the_oracle(cuda):
    """datapath.mix3d_merge (utils/datasets/mix3D.py:43-87) with the CUDA voxelisation == the same function on the
    oracle shim (which tests/test_reference_unchanged.py pins to the reference's own merge_data), bit exact."""
    from lidog_b200.lidog import datapath, synth
    from oracle import me_cpu

    def sample(seed, idx):
        pts, lab = synth.make_scan(seed, "nuscenes")
        q, f, _, vidx, _ = ov.sparse_quantize(pts, np.ones((len(pts), 1), np.float32), lab, -1, True, True, False, 0.05)
        return dict(coordinates=torch.from_numpy(q), xyz=torch.from_numpy(pts[vidx]), features=torch.from_numpy(f),
                    sem_labels=torch.from_numpy(lab[vidx]), sampled_idx=torch.from_numpy(vidx), idx=torch.tensor(idx))
    s0, s1 = sample(41, 0), sample(42, 1)
    ref = datapath.mix3d_merge(s0, s1, 0.05, -1, ME=me_cpu)
    dev = lambda s: {k: v.to(cuda) for k, v in s.items()}
    got = datapath.mix3d_merge(dev(s0), dev(s1), 0.05, -1)
    for k in ("coordinates", "features", "sem_labels", "sampled_idx"):
        assert got[k].is_cuda and torch.equal(got[k].cpu(), torch.as_tensor(ref[k])), k
