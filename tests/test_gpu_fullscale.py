"""GPU parity at BASELINE scale ("no parity test at BASELINE scale").

The bench runs a batch of 8 SemanticKITTI-shaped scans: 1.02 M points, 648 k voxels, 3.97 M pairs in a 3x3x3 map,
~5 000 row tiles, super-tiles of 2-4 tiles with two accumulator sets and a dynamic multi-wave schedule.  These tests
put exactly that batch (and a Mix3D-shaped sample, BASELINE configs[4]) through the CUDA path and the numpy / float64
oracle: voxel sets, maps and pair lists bit-exact, convolution outputs and gradients within 1e-3.
ME conventions behind "bit-exact" are [ME-recalled]: parity is pinned to the oracle, not to
MinkowskiEngine itself, which is absent.
"""
import numpy as np
import pytest
import torch

from oracle import conv as oc
from oracle import voxel as ov
from tests.helpers import record

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kitti_batch():
    from lidog_b200.lidog import synth
    return synth.make_batch(8, 1234, "kitti", 7)  # the bench's rank-0 batch


@pytest.fixture(scope="module")
def kitti_oracle(kitti_batch):
    """Per-scan oracle voxelisation, batched like the collation does (collation.py:309-325)."""
    qs, umaps, invs, cols = [], [], [], []
    off_p = off_v = 0
    for p, l in kitti_batch:
        q, _, cl, um, inv = ov.sparse_quantize(p, np.ones((len(p), 1), np.float32), l, -1, True, True, False, 0.05)
        qs.append(q), cols.append(cl), umaps.append(um + off_p), invs.append(inv + off_v)
        off_p += len(p)
        off_v += len(q)
    coords = ov.batched_coordinates(qs)
    levels = {1: coords}
    parents = {}
    cur = coords
    for ts in (2, 4, 8, 16):
        cur, parents[ts] = ov.stride_coords(cur, ts)
        levels[ts] = cur
    return dict(coords=coords, unique_map=np.concatenate(umaps), inverse_map=np.concatenate(invs),
                colabels=np.concatenate(cols), levels=levels, parents=parents)


@pytest.fixture(scope="module")
def kitti_gpu(cuda, kitti_batch):
    import MinkowskiEngine as ME
    pts = [torch.from_numpy(p).to(cuda) for p, _ in kitti_batch]
    lab = [torch.from_numpy(l).to(cuda) for _, l in kitti_batch]
    q = ME.utils.sparse_quantize_batch(pts, lab, 0.05, -1)
    from lidog_b200.me.coords import CoordinateManager
    return q, CoordinateManager.from_quantized(q)


def test_fullscale_voxelisation_is_bit_exact(kitti_gpu, kitti_oracle):
    q, _ = kitti_gpu
    o = kitti_oracle
    assert q["coords"].shape[0] == o["coords"].shape[0] > 600_000
    assert np.array_equal(q["coords"].cpu().numpy(), o["coords"])
    assert np.array_equal(q["unique_map"].cpu().numpy(), o["unique_map"])
    assert np.array_equal(q["inverse_map"].cpu().numpy(), o["inverse_map"])
    assert np.array_equal(q["colabels"].cpu().numpy(), o["colabels"])
    record("fullscale_voxelisation", points=int(q["inverse_map"].shape[0]), voxels=int(q["coords"].shape[0]))


def test_fullscale_coordinate_levels_are_bit_exact(kitti_gpu, kitti_oracle):
    """All five levels come out of ONE lg_coords_pyramid call (device-side counts between the levels)."""
    _, cm = kitti_gpu
    for ts in (2, 4, 8, 16):
        lvl = cm.levels[ts]  # prebuilt: no further library call
        assert np.array_equal(lvl.coords.cpu().numpy(), kitti_oracle["levels"][ts]), ts
        assert np.array_equal(lvl.parent_of_finer.cpu().numpy(), kitti_oracle["parents"][ts]), ts
    record("fullscale_levels", **{f"ts{ts}": int(cm.levels[ts].n) for ts in (1, 2, 4, 8, 16)})


@pytest.mark.parametrize("ksize,ts", [(3, 1), (3, 16), (5, 1)])
def test_fullscale_kernel_map_pairs_are_bit_exact(kitti_gpu, kitti_oracle, ksize, ts):
    _, cm = kitti_gpu
    c = kitti_oracle["levels"][ts]
    plan = cm.plan("same", ts, ts, ksize)
    i, o, koff = cm.kernel_map_pairs(plan)
    i, o, koff = i.cpu().numpy(), o.cpu().numpy(), koff.numpy()
    ref = ov.kernel_map(c, c, ksize, ts)
    assert koff[-1] == sum(len(m[0]) for m in ref)
    for k, (ri, ro) in enumerate(ref):
        assert np.array_equal(i[koff[k]:koff[k + 1]], ri), k
        assert np.array_equal(o[koff[k]:koff[k + 1]], ro), k
    record("fullscale_kernel_map", ksize=ksize, ts=ts, pairs=int(koff[-1]), rows=int(c.shape[0]))
    cm.plans.pop(("same", ts, ts, ksize))  # the k5 table is 324 MB: do not keep it for the module


def test_fullscale_stride2_and_transposed_maps_are_bit_exact(kitti_gpu, kitti_oracle):
    _, cm = kitti_gpu
    fine, coarse = kitti_oracle["levels"][1], kitti_oracle["levels"][2]
    down = cm.plan("down", 1, 2, 2)
    i, o, koff = (t.cpu().numpy() if hasattr(t, "cpu") else t for t in cm.kernel_map_pairs(down))
    ref = ov.kernel_map(fine, coarse, 2, 1)
    assert koff[-1] == fine.shape[0]
    for k, (ri, ro) in enumerate(ref):
        assert np.array_equal(i[koff[k]:koff[k + 1]], ri) and np.array_equal(o[koff[k]:koff[k + 1]], ro), k
    up = cm.plan("up", 2, 1, 2)
    g, orow = up.nbr.cpu().numpy(), up.out_row.cpu().numpy()
    mask = up.tile_mask.cpu().numpy().view(np.uint32).reshape(-1)
    reft = ov.transposed_kernel_map(fine, coarse, 2, 1)
    tiles_k = np.array([int(m).bit_length() - 1 for m in mask])  # one offset per tile (or -1: unused tile)
    valid = orow >= 0
    for k, (ri, ro) in enumerate(reft):
        sel = valid & np.repeat(tiles_k == k, 128)
        assert np.array_equal(g[sel], ri) and np.array_equal(orow[sel], ro), k


def test_fullscale_sorted_plan_is_a_row_permutation(kitti_gpu):
    _, cm = kitti_gpu
    nat, srt = cm.plan("same", 1, 1, 3), cm.plan("same_sorted", 1, 1, 3)
    n = nat.n_out
    orow = srt.out_row.cpu().numpy()
    assert np.array_equal(np.sort(orow[:n]), np.arange(n)) and np.all(orow[n:] == -1)
    assert torch.equal(srt.nbr[:, :n], nat.nbr[:, srt.out_row[:n].long()])
    units = lambda p: int(sum(bin(int(v)).count("1") for v in p.tile_mask.cpu().numpy().view(np.uint32).reshape(-1)))
    record("fullscale_sorted_plan", units_natural=units(nat), units_sorted=units(srt), tiles=int(srt.n_tiles))


def rel_err(got, ref):
    ref = ref.double()
    return float((got.double().cpu() - ref).norm() / ref.norm().clamp_min(1e-30))


@pytest.mark.parametrize("ts,cin,cout", [(1, 96, 96), (16, 256, 256), (2, 32, 32), (8, 128, 128)])
def test_fullscale_convolution_matches_float64_oracle(cuda, kitti_gpu, kitti_oracle, ts, cin, cout):
    """The PRODUCTION schedule (super-tiles of 2-4 tiles, two accumulator sets, >= 9 dynamically drawn super-tiles per
    CTA at tensor stride 1, many ring wraps) against the float64 oracle: forward, dgrad, wgrad <= 1e-3, and the
    epilogue's batch-norm statistics against the column sums of the result it wrote."""
    import MinkowskiEngine as ME
    _, cm = kitti_gpu
    c = kitti_oracle["levels"][ts]
    n = c.shape[0]
    maps = ov.kernel_map(c, c, 3, ts)
    torch.manual_seed(ts)
    layer = ME.MinkowskiConvolution(cin, cout, kernel_size=3, dimension=3).to(cuda)
    x = torch.randn(n, cin, device=cuda).relu_().requires_grad_(True)
    y = layer(ME.SparseTensor(x, tensor_stride=ts, coordinate_manager=cm))
    gy = torch.randn(n, cout, device=cuda) * 1e-4
    y.F.backward(gy)
    xr = x.detach().cpu().double().requires_grad_(True)
    wr = layer.kernel.detach().cpu().double().requires_grad_(True)
    yr = oc.SparseConvFunction.apply(xr, wr, maps, n)
    yr.backward(gy.cpu().double())
    errs = dict(y=rel_err(y.F.detach(), yr.detach()), dx=rel_err(x.grad, xr.grad), dw=rel_err(layer.kernel.grad, wr.grad))
    sp = y._stat_partials
    if sp is not None:  # epilogue statistics: [4 * tiles, 2 * Cout] partial rows -> column sums / sums of squares
        tot = sp[0].double().sum(0).cpu()
        errs["stat_sum"] = rel_err(tot[:cout], yr.detach().sum(0))
        errs["stat_sq"] = rel_err(tot[cout:], yr.detach().square().sum(0))
    record("fullscale_conv", ts=ts, cin=cin, cout=cout, rows=n, **errs)
    for k, v in errs.items():
        assert v <= 1e-3, (ts, cin, cout, k, v)
    assert sp is not None, "the layer-level forward should have emitted the batch-norm statistics"


def test_fullscale_mix3d_sample_is_bit_exact(cuda):
    """BASELINE configs[4]: two voxelised kitti-shaped scans, back to float32 metres, merged, re-quantised
    (utils/datasets/mix3D.py:43-87) -- the merged cloud holds the VOXEL centres of both scans (~137 k rows from
    ~250 k raw points); the float32 re-quantisation trap is inside."""
    import MinkowskiEngine as ME
    from lidog_b200.lidog import synth
    pts, lab = synth.make_mix3d_scan(77, 7)
    q_ref, _, cl_ref, um_ref, inv_ref = ov.sparse_quantize(pts, np.ones((len(pts), 1), np.float32), lab, -1, True, True,
                                                           False, 0.05)
    q = ME.utils.sparse_quantize_batch([torch.from_numpy(pts).to(cuda)], [torch.from_numpy(lab).to(cuda)], 0.05, -1)
    assert len(pts) > 120_000  # two voxelised kitti-shaped scans (~69 k voxels each after the crop)
    assert np.array_equal(q["coords"].cpu().numpy()[:, 1:], q_ref)
    assert np.array_equal(q["unique_map"].cpu().numpy(), um_ref)
    assert np.array_equal(q["inverse_map"].cpu().numpy(), inv_ref)
    assert np.array_equal(q["colabels"].cpu().numpy(), cl_ref)
    cur = ov.batched_coordinates([q_ref])
    for lv in q["levels"][1:]:
        cur, inv = ov.stride_coords(cur, lv["stride"])
        assert np.array_equal(lv["coords"].cpu().numpy(), cur) and np.array_equal(lv["inverse_map"].cpu().numpy(), inv)
    record("fullscale_mix3d", points=len(pts), voxels=len(q_ref))
