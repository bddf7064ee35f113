"""GPU parity: sparse convolution forward / dgrad / wgrad vs the oracle.

Tolerance (north_star): relative error <= 1e-3 on fp32 outputs and gradients, measured as
||got - ref||_F / ||ref||_F against the oracle evaluated in float64 on the same inputs.
The SIMT fp32 kernels are held to 1e-5.
"""
import numpy as np
import pytest
import torch

from oracle import voxel as ov
from oracle import conv as oc
from tests.helpers import random_voxels

pytestmark = pytest.mark.gpu

TOL_TC = 1e-3
TOL_SIMT = 1e-5


def rel_err(got, ref):
    ref = ref.double()
    return float((got.double().cpu() - ref).norm() / ref.norm().clamp_min(1e-30))


def _oracle_maps(cm_coords, kind, ts_in, ksize):
    if kind == "same":
        c = cm_coords[ts_in]
        return ov.kernel_map(c, c, ksize, ts_in), c.shape[0], c.shape[0]
    if kind == "down":
        return ov.kernel_map(cm_coords[ts_in], cm_coords[2 * ts_in], 2, ts_in), cm_coords[ts_in].shape[0], cm_coords[2 * ts_in].shape[0]
    if kind == "up":
        return (ov.transposed_kernel_map(cm_coords[ts_in // 2], cm_coords[ts_in], 2, ts_in // 2),
                cm_coords[ts_in].shape[0], cm_coords[ts_in // 2].shape[0])
    if kind == "identity":
        n = cm_coords[ts_in].shape[0]
        r = np.arange(n)
        return [(r, r)], n, n


def _run_case(cuda, mode, kind, ts_in, ksize, cin, cout, bias, tol, seed=0, sorted_plans=1, n_vox=9000):
    import MinkowskiEngine as ME
    from lidog_b200.me import conv as meconv
    rng = np.random.default_rng(seed)
    coords = random_voxels(rng, n_vox, span=30)
    old = dict(meconv.CONFIG)
    meconv.CONFIG["tc"] = mode
    meconv.CONFIG["sorted"] = sorted_plans
    try:
        feats1 = torch.ones(coords.shape[0], 1, device=cuda)
        base = ME.SparseTensor(coordinates=torch.from_numpy(coords).to(cuda), features=feats1)
        cm = base.coordinate_manager
        cm_coords = {ts: cm.get_coords(ts).cpu().numpy() for ts in (1, 2, 4)}
        maps, n_in, n_out = _oracle_maps(cm_coords, kind, ts_in, ksize)
        torch.manual_seed(seed)
        cls = ME.MinkowskiConvolutionTranspose if kind == "up" else ME.MinkowskiConvolution
        layer = cls(cin, cout, kernel_size=ksize, stride=2 if kind in ("down", "up") else 1, bias=bias,
                    dimension=3).to(cuda)
        x = torch.randn(n_in, cin, device=cuda).relu_().requires_grad_(True)  # post-ReLU-like operand
        xin = ME.SparseTensor(x, tensor_stride=ts_in, coordinate_manager=cm)
        y = layer(xin)
        assert y.F.shape == (n_out, cout)
        gy = torch.randn(n_out, cout, device=cuda) * 1e-4  # small gradients exercise the fp16 scaling
        y.F.backward(gy)
        # oracle in float64
        xr = x.detach().cpu().double().requires_grad_(True)
        wr = layer.kernel.detach().cpu().double().requires_grad_(True)
        yr = oc.SparseConvFunction.apply(xr, wr, maps, n_out)
        if bias:
            yr = yr + layer.bias.detach().cpu().double()
        yr.backward(gy.cpu().double())
        errs = dict(y=rel_err(y.F.detach(), yr.detach()), dx=rel_err(x.grad, xr.grad),
                    dw=rel_err(layer.kernel.grad, wr.grad))
        if bias:
            errs["db"] = rel_err(layer.bias.grad, gy.cpu().double().sum(0, keepdim=True))
        for k, v in errs.items():
            assert v <= tol, (mode, kind, cin, cout, k, v)
        return errs
    finally:
        meconv.CONFIG.update(old)


SIMT_CASES = [
    ("same", 1, 5, 1, 32, False),    # stem conv0p1s1
    ("identity", 1, 1, 96, 7, True),  # class head with bias (narrow-head kernels, csrc/conv_simt.cu)
    ("identity", 1, 1, 256, 3, False),  # narrow head, widest operand, no bias
    ("identity", 2, 1, 32, 8, True),    # narrow head, full 8 columns
    ("identity", 1, 1, 96, 19, True),  # 19-class head (BASELINE configs[0]): generic tile kernel
    ("same", 1, 3, 32, 32, False),
    ("same", 2, 3, 32, 64, False),
    ("down", 1, 2, 32, 32, False),
    ("up", 2, 2, 64, 48, False),
    ("identity", 2, 1, 24, 40, False),
]


@pytest.mark.parametrize("kind,ts_in,ksize,cin,cout,bias", SIMT_CASES)
def test_conv_simt_matches_oracle(cuda, kind, ts_in, ksize, cin, cout, bias):
    _run_case(cuda, "off", kind, ts_in, ksize, cin, cout, bias, TOL_SIMT)


TC_CASES = [
    ("same", 1, 3, 32, 32), ("same", 1, 3, 64, 64), ("same", 2, 3, 32, 64), ("same", 1, 3, 128, 96),
    ("same", 1, 3, 96, 96), ("same", 2, 3, 64, 128), ("same", 4, 3, 128, 256), ("same", 4, 3, 256, 256),
    ("same", 2, 3, 384, 256), ("same", 2, 3, 192, 128),
    ("down", 1, 2, 32, 32), ("down", 2, 2, 64, 64), ("down", 1, 2, 128, 128),
    ("up", 2, 2, 256, 256), ("up", 2, 2, 256, 128), ("up", 4, 2, 128, 96), ("up", 2, 2, 96, 96),
    ("identity", 1, 1, 32, 64), ("identity", 2, 1, 384, 256), ("identity", 1, 1, 128, 96),
]


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
@pytest.mark.parametrize("kind,ts_in,ksize,cin,cout", TC_CASES)
def test_conv_tc_matches_oracle(cuda, mode, kind, ts_in, ksize, cin, cout):
    # bf16 operands (8 mantissa bits) cannot reach 1e-3; they are held to 4e-3 and reported in DESIGN.md
    tol = TOL_TC if mode == "fp16" else 4e-3
    _run_case(cuda, mode, kind, ts_in, ksize, cin, cout, False, tol)


@pytest.mark.parametrize("kind,ts_in,ksize,cin,cout", [("same", 1, 3, 96, 96), ("same", 2, 3, 64, 128),
                                                       ("down", 1, 2, 32, 32), ("up", 2, 2, 96, 96)])
def test_conv_tc_natural_order_plans(cuda, kind, ts_in, ksize, cin, cout):
    """LIDOG_SORTED_PLANS=0: the natural-row-order plans stay a supported configuration."""
    _run_case(cuda, "fp16", kind, ts_in, ksize, cin, cout, False, TOL_TC, sorted_plans=0)


@pytest.mark.parametrize("n_vox", [1, 100, 127, 129, 40000])
def test_conv_tc_ragged_sizes(cuda, n_vox):
    """Tile-boundary and tiny inputs (1 voxel, < 1 tile, 1 tile + 1) and a multi-super-tile case."""
    _run_case(cuda, "fp16", "same", 1, 3, 64, 64, False, TOL_TC, seed=n_vox, n_vox=n_vox)


def test_conv_matches_dense_conv3d(cuda):
    """ME-independent ground truth: densify a small grid and use torch.nn.functional.conv3d."""
    import MinkowskiEngine as ME
    from lidog_b200.me import conv as meconv
    rng = np.random.default_rng(7)
    G = 12
    occ = rng.random((G, G, G)) < 0.3
    xyz = np.argwhere(occ).astype(np.int32)
    coords = np.concatenate([np.zeros((len(xyz), 1), np.int32), xyz], 1)
    cin, cout = 32, 32
    old = dict(meconv.CONFIG)
    try:
        for mode, tol in (("off", 1e-5), ("fp16", 1e-3)):
            meconv.CONFIG["tc"] = mode
            torch.manual_seed(0)
            x = torch.randn(len(xyz), cin, device=cuda)
            layer = ME.MinkowskiConvolution(cin, cout, kernel_size=3, dimension=3).to(cuda)
            y = layer(ME.SparseTensor(coordinates=torch.from_numpy(coords).to(cuda), features=x)).F
            dense = torch.zeros(1, cin, G, G, G, dtype=torch.float64)
            dense[0, :, xyz[:, 2], xyz[:, 1], xyz[:, 0]] = x.cpu().double().t()  # (z, y, x) spatial order
            # kernel index k = ix + 3*(iy + 3*iz)  ->  weight[co, ci, iz, iy, ix]
            w = layer.kernel.detach().cpu().double().view(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2)
            ref = torch.nn.functional.conv3d(dense, w, padding=1)[0][:, xyz[:, 2], xyz[:, 1], xyz[:, 0]].t()
            assert rel_err(y.detach(), ref) <= tol, mode
    finally:
        meconv.CONFIG.update(old)
