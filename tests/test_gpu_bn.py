"""GPU parity: fused BN (+ ReLU, residual, second BN branch) vs torch BatchNorm1d / relu / add in float64.

Reference contract: ME.MinkowskiBatchNorm is torch BatchNorm1d on the feature matrix, chained with
MinkowskiReLU and BasicBlock's residual add (utils/models/minkunet_bev.py:308-368).  The checker is the same
module chain evaluated by torch in float64 (the unfused path of the shim): torch's own float32 channels-last
BN backward is off by 2.5e-3 at C = 256 on this input (measured, tools/bn_dbg.py), so it cannot be the yardstick.
Tolerance: 1e-5 relative (Frobenius) on outputs, gradients and running statistics."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    b = b.double()
    return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))


def _sparse(cuda, n):
    import MinkowskiEngine as ME
    from tests.helpers import random_voxels
    coords = random_voxels(np.random.default_rng(3), n, span=40)
    base = ME.SparseTensor(coordinates=torch.from_numpy(coords).to(cuda),
                           features=torch.ones(coords.shape[0], 1, device=cuda))
    return ME, base.coordinate_manager, coords.shape[0]


@pytest.fixture(params=[1, 0], ids=["layer_calls", "fine_calls"])
def layer_calls(request):
    """Both host paths: one library call per layer each way (default) and the fine-grained entry points."""
    from lidog_b200.me import norm, conv
    old = norm.CONFIG["layer_calls"], conv.CONFIG["layer_calls"]
    norm.CONFIG["layer_calls"] = conv.CONFIG["layer_calls"] = request.param
    yield request.param
    norm.CONFIG["layer_calls"], conv.CONFIG["layer_calls"] = old


@pytest.mark.parametrize("C", [32, 96, 256])
@pytest.mark.parametrize("mode", ["bn_relu", "bn", "bn_res_relu", "bn_bn_relu", "bn_relu_cat"])
def test_fused_bn_matches_torch(cuda, C, mode, layer_calls):
    ME, cm, n = _sparse(cuda, 5000)
    torch.manual_seed(C)
    x = (torch.randn(n, C, device=cuda) * 2 + 0.5).requires_grad_(True)
    x2 = (torch.randn(n, C, device=cuda) - 1.0).requires_grad_(True)
    res = torch.randn(n, C, device=cuda).requires_grad_(True)
    gy = torch.randn(n, C, device=cuda) * 1e-3

    def make():
        a, b = ME.MinkowskiBatchNorm(C).to(cuda), ME.MinkowskiBatchNorm(C).to(cuda)
        with torch.no_grad():
            for m, s in ((a, 1), (b, 2)):
                g = torch.Generator(device="cpu").manual_seed(s)
                m.bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
                m.bn.bias.copy_(torch.randn(C, generator=g) * 0.1)
        return a, b

    def run(fused):
        from lidog_b200.me import norm
        old = norm.CONFIG["fused"]
        norm.CONFIG["fused"] = fused
        dt = torch.float32 if fused else torch.float64
        try:
            a, b = make()
            a, b = a.to(dt), b.to(dt)
            relu = ME.MinkowskiReLU(inplace=True)
            xs = [t.detach().clone().to(dt).requires_grad_(True) for t in (x, x2, res)]
            out = a(ME.SparseTensor(xs[0], coordinate_manager=cm))
            if mode == "bn_res_relu":
                out += ME.SparseTensor(xs[2], coordinate_manager=cm)
            if mode == "bn_bn_relu":
                out += b(ME.SparseTensor(xs[1], coordinate_manager=cm))
            if mode != "bn":
                out = relu(out)
            y = out.F
            if mode == "bn_relu_cat":  # decoder: ME.cat(out, skip) -- the gradient comes back as a column slice
                z = torch.cat([y, xs[2]], dim=1)
                z.backward(torch.cat([gy, gy * 0.5], dim=1).to(dt))
            else:
                y.backward(gy.to(dt))
            grads = [t.grad for t in xs] + [a.bn.weight.grad, a.bn.bias.grad, b.bn.weight.grad, b.bn.bias.grad]
            stats = [a.bn.running_mean, a.bn.running_var, a.bn.num_batches_tracked]
            return y.detach(), grads, stats, out
        finally:
            norm.CONFIG["fused"] = old

    y_f, g_f, s_f, out_f = run(1)
    y_t, g_t, s_t, _ = run(0)
    assert rel(y_f, y_t) <= 1e-5
    for gf, gt in zip(g_f, g_t):
        assert (gf is None) == (gt is None)
        if gt is not None:
            assert rel(gf, gt) <= 2e-5
    assert rel(s_f[0], s_t[0]) <= 1e-5 and rel(s_f[1], s_t[1]) <= 1e-5 and int(s_f[2]) == int(s_t[2]) == 1
    # the 16-bit operand copy rides along and equals the cast of y
    from lidog_b200 import cabi
    y16 = out_f._f16(cabi.FMT_FP16)
    assert y16.dtype == torch.float16 and torch.equal(y16, y_f.clamp(-65504, 65504).half())


def test_fused_bn_hands_scaled_fp16_gradient_to_the_convolution(cuda, layer_calls):
    """conv -> BN -> ReLU: the BN backward publishes dx as a scaled fp16 copy; the convolution's backward uses it
    and the result equals the unfused path within the tensor-core tolerance."""
    ME, cm, n = _sparse(cuda, 6000)
    from lidog_b200.me import norm, _grad16
    torch.manual_seed(0)
    conv = ME.MinkowskiConvolution(64, 64, kernel_size=3, dimension=3).to(cuda)
    x0 = torch.randn(n, 64, device=cuda).relu_()
    gy = torch.randn(n, 64, device=cuda) * 1e-4
    outs = []
    for fused in (1, 0):
        norm.CONFIG["fused"] = fused
        try:
            bn = ME.MinkowskiBatchNorm(64).to(cuda)
            x = x0.clone().requires_grad_(True)
            conv.kernel.grad = None
            y = ME.MinkowskiReLU()(bn(conv(ME.SparseTensor(x, coordinate_manager=cm)))).F
            y.backward(gy)
            outs.append((y.detach(), x.grad.clone(), conv.kernel.grad.clone()))
            assert len(_grad16._TABLE) == 0  # the published copy was consumed (fused) / never made (unfused)
        finally:
            norm.CONFIG["fused"] = 1
    for a, b in zip(outs[0], outs[1]):
        assert rel(a, b) <= 1e-3


@pytest.mark.parametrize("C", [32, 96, 256])
def test_epilogue_statistics_feed_the_batch_norm(cuda, C):
    """conv -> BN -> ReLU with the statistics taken from the GEMM epilogue (LIDOG_EPI_STATS=1, the default) equals
    the same chain with the separate statistics pass (=0): outputs, every gradient, running statistics."""
    ME, cm, n = _sparse(cuda, 7000)
    from lidog_b200.me import conv as meconv
    torch.manual_seed(1)
    conv = ME.MinkowskiConvolution(64, C, kernel_size=3, dimension=3).to(cuda)
    x0 = torch.randn(n, 64, device=cuda).relu_()
    gy = torch.randn(n, C, device=cuda) * 1e-3
    outs = []
    old = meconv.CONFIG["epi_stats"]
    try:
        for epi in (1, 0):
            meconv.CONFIG["epi_stats"] = epi
            bn = ME.MinkowskiBatchNorm(C).to(cuda)
            x = x0.clone().requires_grad_(True)
            conv.kernel.grad = None
            c = conv(ME.SparseTensor(x, coordinate_manager=cm))
            assert (c._stat_partials is not None) == bool(epi)
            y = ME.MinkowskiReLU()(bn(c)).F
            y.backward(gy)
            outs.append((y.detach(), x.grad.clone(), conv.kernel.grad.clone(), bn.bn.weight.grad.clone(),
                         bn.bn.bias.grad.clone(), bn.bn.running_mean.clone(), bn.bn.running_var.clone()))
    finally:
        meconv.CONFIG["epi_stats"] = old
    # (the two convolution gradients pass through the fp16 rounding of the BN's dx: a 1e-7 change of a value flips
    # the rounding of a few elements, which is worth ~1e-5 in the Frobenius norm)
    for (a, b), tol in zip(zip(outs[0], outs[1]), (1e-5, 1e-4, 1e-4, 1e-5, 1e-5, 1e-5, 1e-5)):
        assert rel(a, b) <= tol


def test_fused_bn_eval_mode_uses_running_statistics(cuda):
    ME, cm, n = _sparse(cuda, 2000)
    bn = ME.MinkowskiBatchNorm(32).to(cuda)
    x = torch.randn(n, 32, device=cuda)
    bn.train()
    ME.MinkowskiReLU()(bn(ME.SparseTensor(x, coordinate_manager=cm)))
    bn.eval()
    y = bn(ME.SparseTensor(x, coordinate_manager=cm)).F
    ref = torch.nn.functional.batch_norm(x, bn.bn.running_mean, bn.bn.running_var, bn.bn.weight, bn.bn.bias, False)
    assert rel(y, ref) <= 1e-6
