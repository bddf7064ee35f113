"""CPU: the oracle's sparse convolution against dense torch conv3d (ME-independent ground truth)
and fp64 gradcheck of its autograd wrapper."""
import numpy as np
import torch

from oracle import voxel as ov
from oracle import conv as oc


def _grid(rng, G=8, p=0.35):
    xyz = np.argwhere(rng.random((G, G, G)) < p).astype(np.int32)
    return xyz, np.concatenate([np.zeros((len(xyz), 1), np.int32), xyz], 1)


def test_k3_matches_dense_conv3d():
    rng = np.random.default_rng(0)
    G = 8
    xyz, coords = _grid(rng, G)
    cin, cout = 5, 7
    x = torch.randn(len(xyz), cin, dtype=torch.float64)
    w = torch.randn(27, cin, cout, dtype=torch.float64)
    y = oc.conv_forward(x, w, ov.kernel_map(coords, coords, 3, 1), len(xyz))
    dense = torch.zeros(1, cin, G, G, G, dtype=torch.float64)
    dense[0, :, xyz[:, 2], xyz[:, 1], xyz[:, 0]] = x.t()
    wd = w.view(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2)  # k = ix + 3*(iy + 3*iz)
    ref = torch.nn.functional.conv3d(dense, wd, padding=1)[0][:, xyz[:, 2], xyz[:, 1], xyz[:, 0]].t()
    assert torch.allclose(y, ref, atol=1e-10)


def test_k2s2_and_transpose_match_dense():
    rng = np.random.default_rng(1)
    G = 8
    xyz, coords = _grid(rng, G, 0.5)
    c2, _ = ov.stride_coords(coords, 2)
    cin, cout = 3, 4
    x = torch.randn(len(xyz), cin, dtype=torch.float64)
    w = torch.randn(8, cin, cout, dtype=torch.float64)
    y = oc.conv_forward(x, w, ov.kernel_map(coords, c2, 2, 1), len(c2))
    dense = torch.zeros(1, cin, G, G, G, dtype=torch.float64)
    dense[0, :, xyz[:, 2], xyz[:, 1], xyz[:, 0]] = x.t()
    wd = w.view(2, 2, 2, cin, cout).permute(4, 3, 0, 1, 2)
    ref = torch.nn.functional.conv3d(dense, wd, stride=2)[0]
    got = ref[:, c2[:, 3] // 2, c2[:, 2] // 2, c2[:, 1] // 2].t()
    assert torch.allclose(y, got, atol=1e-10)
    # transposed: coarse -> the existing fine set
    xc = torch.randn(len(c2), cout, dtype=torch.float64)
    wt = torch.randn(8, cout, cin, dtype=torch.float64)
    yt = oc.conv_forward(xc, wt, ov.transposed_kernel_map(coords, c2, 2, 1), len(xyz))
    densec = torch.zeros(1, cout, G // 2, G // 2, G // 2, dtype=torch.float64)
    densec[0, :, c2[:, 3] // 2, c2[:, 2] // 2, c2[:, 1] // 2] = xc.t()
    wtd = wt.view(2, 2, 2, cout, cin).permute(3, 4, 0, 1, 2)  # conv_transpose3d weight: (in, out, kz, ky, kx)
    reft = torch.nn.functional.conv_transpose3d(densec, wtd, stride=2)[0][:, xyz[:, 2], xyz[:, 1], xyz[:, 0]].t()
    assert torch.allclose(yt, reft, atol=1e-10)


def test_autograd_wrapper_gradcheck():
    rng = np.random.default_rng(2)
    xyz, coords = _grid(rng, 5, 0.4)
    maps = [(torch.from_numpy(i), torch.from_numpy(o)) for i, o in ov.kernel_map(coords, coords, 3, 1)]
    x = torch.randn(len(xyz), 3, dtype=torch.float64, requires_grad=True)
    w = torch.randn(27, 3, 2, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: oc.SparseConvFunction.apply(a, b, maps, len(xyz)), (x, w))
    # explicit dgrad / wgrad agree with autograd of the forward
    y = oc.conv_forward(x, w, maps, len(xyz))
    gy = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, (x, w), gy)
    assert torch.allclose(gx, oc.conv_dgrad(gy, w.detach(), maps, len(xyz)))
    assert torch.allclose(gw, oc.conv_wgrad(x.detach(), gy, maps, 27))
