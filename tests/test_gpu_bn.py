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


@pytest.mark.parametrize("shape", [(2, 64, 37, 41), (3, 256, 20, 23)])
def test_dense_head_bn_relu_on_the_fused_kernels(cuda, shape):
    """BatchNorm2d -> ReLU of the dense BEV head (utils/models/conv2d.py:9-25) through bn_relu_2d: a channels_last
    image is a [B*H*W, C] matrix.  Output, dx, dgamma, dbeta, running statistics vs torch in float64; an NCHW input
    takes the torch modules (and must give the same numbers)."""
    from lidog_b200 import cabi
    from lidog_b200.me import norm
    from tests.helpers import record
    torch.manual_seed(5)
    B, C, H, W = shape
    x0 = (torch.randn(B, C, H, W, device=cuda) * 2 + 0.3)
    gy = torch.randn(B, C, H, W, device=cuda)
    ref_bn = torch.nn.BatchNorm2d(C).to(cuda).double()
    with torch.no_grad():
        ref_bn.weight.uniform_(0.5, 1.5)
        ref_bn.bias.uniform_(-0.5, 0.5)
    xr = x0.double().requires_grad_(True)
    yr = torch.relu(ref_bn(xr))
    yr.backward(gy.double())
    for fmt_name, mem in (("channels_last", torch.channels_last), ("nchw", torch.contiguous_format)):
        bn = torch.nn.BatchNorm2d(C).to(cuda)
        with torch.no_grad():
            bn.weight.copy_(ref_bn.weight)
            bn.bias.copy_(ref_bn.bias)
        x = x0.clone().contiguous(memory_format=mem).requires_grad_(True)
        cabi.counting(True)  # census mode of the binding: which library entry points ran
        try:
            y = norm.bn_relu_2d(x, bn)
            y.backward(gy.contiguous(memory_format=mem))
            fused = cabi.COUNTS.get("#lg_bn_layer_forward", 0) > 0 and cabi.COUNTS.get("lg_bn_layer_backward", 0) == 1
        finally:
            cabi.counting(False)
        assert fused == (fmt_name == "channels_last")
        errs = {"y": rel(y, yr), "dx": rel(x.grad, xr.grad), "dgamma": rel(bn.weight.grad, ref_bn.weight.grad),
                "dbeta": rel(bn.bias.grad, ref_bn.bias.grad), "running_mean": rel(bn.running_mean, ref_bn.running_mean),
                "running_var": rel(bn.running_var, ref_bn.running_var)}
        record("dense_head_bn_relu", shape=list(shape), memory=fmt_name, **errs)
        bar = 1e-5 if fused else 5e-3  # torch's own float32 BN backward is the loose one (module docstring)
        assert all(v <= bar for v in errs.values()), (fmt_name, errs)
        assert int(bn.num_batches_tracked) == 1


def test_skipped_fp32_gradient_is_guarded(cuda):
    """conv -> BN -> ReLU: the BN backward hands the convolution a 16-bit gradient only (LIDOG_BN_SKIP_DX32=1).  Same
    results as with the fp32 gradient written; a second consumer of the convolution output must fail loudly."""
    ME, cm, n = _sparse(cuda, 5000)
    from lidog_b200.me import norm, _grad16
    torch.manual_seed(2)
    conv = ME.MinkowskiConvolution(32, 64, kernel_size=3, dimension=3).to(cuda)
    x0 = torch.randn(n, 32, device=cuda).relu_()
    gy = torch.randn(n, 64, device=cuda) * 1e-3
    outs = []
    for skip in (1, 0):
        norm.CONFIG["skip_dx32"] = skip
        try:
            bn = ME.MinkowskiBatchNorm(64).to(cuda)
            x = x0.clone().requires_grad_(True)
            conv.kernel.grad = None
            y = ME.MinkowskiReLU()(bn(conv(ME.SparseTensor(x, coordinate_manager=cm)))).F
            y.backward(gy)
            outs.append((y.detach(), x.grad.clone(), conv.kernel.grad.clone(), bn.bn.weight.grad.clone()))
            assert len(_grad16._TABLE) == 0 and len(_grad16._NO_FP32) == 0
        finally:
            norm.CONFIG["skip_dx32"] = 1
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)  # the 16-bit copy is what the convolution reads either way
    bn = ME.MinkowskiBatchNorm(64).to(cuda)
    mid = conv(ME.SparseTensor(x0.clone().requires_grad_(True), coordinate_manager=cm))
    y = ME.MinkowskiReLU()(bn(mid)).F + mid.F  # second consumer of the convolution output
    with pytest.raises(RuntimeError, match="LIDOG_BN_SKIP_DX32"):
        y.backward(gy)
    _grad16.clear()
