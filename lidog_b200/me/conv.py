"""MinkowskiConvolution / MinkowskiConvolutionTranspose on liblidog_b200.

Mirror of the reference-facing layer interface (ctor arguments, `.kernel`
(K, Cin, Cout) / (Cin, Cout), `.bias` (1, Cout); utils/models/minkunet_bev.py:57-123,
:404).  Forward, dgrad and wgrad call the C ABI: the tcgen05 gather-GEMM kernels for
channel counts that are multiples of 32, the exact-fp32 SIMT kernels otherwise
(1-channel stem, class head).  There is no CPU or torch fallback.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from .. import cabi
from ._grad16 import take_grad16
from .sparse_tensor import SparseTensor

# operand format of the tensor-core path: "fp16" (default; 2^-11 unit round-off meets the 1e-3 bar),
# "bf16", or "off" (SIMT fp32 everywhere).  TC_GATHER: 2 = super-tile pipeline with cp.async row gathers
# arriving on mbarriers (default, csrc/conv_tc2.cu); 0 = first-generation cp.async; 1 = TMA tile::gather4
# (kept for the record: measured 2.7x slower than cp.async on B200, profiles/).
CONFIG = {
    "tc": os.environ.get("LIDOG_TC", "fp16"),
    "gather": int(os.environ.get("LIDOG_TC_GATHER", "2")),
    # mask-sorted gather plans (lg_kernel_map_sorted) for the 3x3x3 and stride-2 layers; 0 = natural row order
    "sorted": int(os.environ.get("LIDOG_SORTED_PLANS", "1")),
}


# bench.py instrumentation: per-launch CUDA events + algorithmic FLOPs (2 * pairs * Cin * Cout).
# A census pass (events=False) counts the pairs of every plan once; timed passes look them up by key.
PROFILE = {"enabled": False, "events": False, "records": [], "pairs": {}}


def _profiled(kernel, plan, cin, cout, launch):
    if not PROFILE["enabled"]:
        return launch()
    if not PROFILE["events"]:
        if plan.key not in PROFILE["pairs"]:
            PROFILE["pairs"][plan.key] = plan.count_pairs()
        out = launch()
        PROFILE["records"].append(dict(kernel=kernel, key=plan.key, cin=cin, cout=cout,
                                       flops=2.0 * PROFILE["pairs"][plan.key] * cin * cout))
        return out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = launch()
    e1.record()
    PROFILE["records"].append(dict(kernel=kernel, key=plan.key, cin=cin, cout=cout, e0=e0, e1=e1,
                                   flops=2.0 * PROFILE["pairs"].get(plan.key, 0) * cin * cout))
    return out


def _fmt():
    return {"fp16": cabi.FMT_FP16, "bf16": cabi.FMT_BF16}.get(CONFIG["tc"])


def _tc_ok(cin, cout):
    return _fmt() is not None and cin % 32 == 0 and cout % 32 == 0


def _cast16(x: torch.Tensor, fmt: int, scale: torch.Tensor | None = None) -> torch.Tensor:
    out = torch.empty(x.shape, dtype=torch.float16 if fmt == cabi.FMT_FP16 else torch.bfloat16, device=x.device)
    cabi.check(cabi.lib().lg_cast_rows(cabi.ptr(x), cabi.ptr(out), x.numel(), fmt, cabi.ptr(scale), cabi.stream()),
               "lg_cast_rows")
    return out


def _gemm_simt(plan, A, W3, N, w_transposed, flip, bias):
    Y = torch.empty((plan.n_out, N), dtype=torch.float32, device=A.device)
    cabi.check(cabi.lib().lg_conv_gemm_simt(plan.c, cabi.ptr(A), A.shape[1], cabi.ptr(W3), N, w_transposed, flip,
                                            cabi.ptr(bias), cabi.ptr(Y), cabi.stream()), "lg_conv_gemm_simt")
    return Y


def _gemm_tc(plan, A16, B16, N, flip, fmt, out_scale, bias):
    Y = torch.empty((plan.n_out, N), dtype=torch.float32, device=A16.device)

    def launch():
        cabi.check(cabi.lib().lg_conv_gemm_tc(plan.c, cabi.ptr(A16), A16.shape[1], cabi.ptr(B16), N, flip, fmt,
                                              cabi.ptr(out_scale), cabi.ptr(bias), cabi.ptr(Y), CONFIG["gather"],
                                              cabi.stream()), "lg_conv_gemm_tc")
    _profiled("k_gemm_tc", plan, A16.shape[1], N, launch)
    return Y


class SparseConvFunction(torch.autograd.Function):
    """y = conv(x, kernel) over the gather plans of one layer.

    plans = (fwd, dgrad, wgrad, flip_dgrad): see MinkowskiConvolutionBase._plans.
    """

    @staticmethod
    def forward(ctx, x, kernel, bias, plans, x16_cache):
        p_fwd = plans[0]
        W3 = kernel if kernel.dim() == 3 else kernel.unsqueeze(0)
        K, cin, cout = W3.shape
        x = x.contiguous()
        W3c = W3.detach().contiguous()
        b = None if bias is None else bias.detach().reshape(-1).contiguous()
        ctx.plans, ctx.shape, ctx.kdim = plans, (K, cin, cout), kernel.dim()
        ctx.has_bias = bias is not None
        ctx.tc = _tc_ok(cin, cout) and x.shape[0] > 0
        if ctx.tc:
            fmt = _fmt()
            x16 = x16_cache(fmt) if x16_cache is not None else _cast16(x.detach(), fmt)
            w16 = torch.empty((K, cin, cout), dtype=x16.dtype, device=x.device)
            w16t = torch.empty((K, cout, cin), dtype=x16.dtype, device=x.device)
            cabi.check(cabi.lib().lg_prep_weights(cabi.ptr(W3c), K, cin, cout, cabi.ptr(w16), cabi.ptr(w16t), fmt,
                                                  cabi.stream()), "lg_prep_weights")
            y = _gemm_tc(p_fwd, x16, w16t, cout, 0, fmt, None, b)
            ctx.fmt, ctx.w16 = fmt, w16
            ctx.save_for_backward(x16)
        else:
            y = _gemm_simt(p_fwd, x.detach(), W3c, cout, 0, 0, b)
            ctx.save_for_backward(x.detach(), W3c)
        return y

    @staticmethod
    def backward(ctx, dy):
        _, p_dgrad, p_wgrad, flip = ctx.plans
        K, cin, cout = ctx.shape
        L = cabi.lib()
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0, keepdim=True)
        if ctx.tc:
            (x16,) = ctx.saved_tensors
            fmt = ctx.fmt
            hit = take_grad16(dy, fmt)  # the fused BN backward already wrote the scaled 16-bit gradient
            if hit is not None:
                dy16, scale = hit
            else:
                scale = None
                if fmt == cabi.FMT_FP16:  # bring the gradient into fp16's normal range (power-of-two scale)
                    scale = torch.empty(4, dtype=torch.float32, device=dy.device)
                    cabi.check(L.lg_absmax_scale(cabi.ptr(dy), dy.numel(), cabi.ptr(scale), cabi.stream()),
                               "lg_absmax_scale")
                dy16 = _cast16(dy, fmt, scale)
            inv = None if scale is None else scale[1:]
            if ctx.needs_input_grad[0]:
                dx = _gemm_tc(p_dgrad, dy16, ctx.w16, cin, flip, fmt, inv, None)
            if ctx.needs_input_grad[1]:
                dw = torch.empty((K, cin, cout), dtype=torch.float32, device=dy.device)
                ws_bytes = L.lg_conv_wgrad_tc_workspace(p_wgrad.c, cin, cout)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dy.device)
                _profiled("k_wgrad_tc", p_wgrad, cin, cout, lambda: cabi.check(
                    L.lg_conv_wgrad_tc(p_wgrad.c, cabi.ptr(x16), cin, cabi.ptr(dy16), cout, fmt, cabi.ptr(inv),
                                       cabi.ptr(dw), CONFIG["gather"], cabi.ptr(ws), ws_bytes, cabi.stream()),
                    "lg_conv_wgrad_tc"))
        else:
            x, W3c = ctx.saved_tensors
            if ctx.needs_input_grad[0]:
                dx = _gemm_simt(p_dgrad, dy, W3c, cin, 1, flip, None)
            if ctx.needs_input_grad[1]:
                dw = torch.empty((K, cin, cout), dtype=torch.float32, device=dy.device)
                ws_bytes = L.lg_conv_wgrad_workspace(p_wgrad.c, cin, cout)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dy.device)
                cabi.check(L.lg_conv_wgrad_simt(p_wgrad.c, cabi.ptr(x), cin, cabi.ptr(dy), cout, cabi.ptr(dw),
                                                cabi.ptr(ws), ws_bytes, cabi.stream()), "lg_conv_wgrad_simt")
        if dw is not None and ctx.kdim == 2:
            dw = dw[0]
        return dx, dw, db, None, None


class MinkowskiConvolutionBase(nn.Module):
    is_transpose = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=3):
        super().__init__()
        if dimension != 3:
            raise ValueError("only dimension=3 is on the LiDOG path")
        if dilation != 1 or kernel_generator is not None or expand_coordinates:
            raise NotImplementedError("dilation / kernel_generator / expand_coordinates are outside the LiDOG path")
        kernel_size, stride = int(kernel_size), int(stride)
        if (kernel_size, stride) not in ((1, 1), (3, 1), (5, 1), (2, 2)):
            raise NotImplementedError(f"kernel_size={kernel_size}, stride={stride} is outside the LiDOG path")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation, self.dimension = kernel_size, stride, 1, dimension
        self.kernel_volume = kernel_size ** 3
        shape = (in_channels, out_channels) if self.kernel_volume == 1 else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self, is_transpose=None):
        # ME default: uniform(+-1/sqrt(fan)), fan = Cin*K (Cout*K when transposed)  [SURVEY App. C.9]
        n = (self.out_channels if self.is_transpose else self.in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def _plans(self, cm, ts_in):
        k, s = self.kernel_size, self.stride
        if self.is_transpose:
            ts_out = ts_in // s
            if s != 2 or ts_in % 2:
                raise NotImplementedError("transposed convolution is implemented for kernel 2 / stride 2")
            if ts_out not in cm.levels:
                raise RuntimeError("transposed convolution needs the finer coordinate map created by the encoder")
            up, down = cm.plan("up", ts_in, ts_out, 2), cm.plan(self._kind("down"), ts_out, ts_in, 2)
            return ts_out, (up, down, up, 0)
        ts_out = ts_in * s
        if s == 2:
            down, up = cm.plan(self._kind("down"), ts_in, ts_out, 2), cm.plan("up", ts_out, ts_in, 2)
            return ts_out, (down, up, down, 0)
        if k == 1:
            p = cm.plan("identity", ts_in, ts_in, 1)
            return ts_out, (p, p, p, 0)
        p = cm.plan(self._kind("same") if k == 3 else "same", ts_in, ts_in, k)
        return ts_out, (p, p, p, 1)

    @staticmethod
    def _kind(kind):
        return kind + "_sorted" if CONFIG["sorted"] else kind

    def forward(self, input: SparseTensor) -> SparseTensor:
        assert isinstance(input, SparseTensor)
        cm = input.coordinate_manager
        ts_out, plans = self._plans(cm, input._ts)
        y = SparseConvFunction.apply(input.F, self.kernel, self.bias, plans, input._f16)
        return SparseTensor(y, tensor_stride=ts_out, coordinate_manager=cm)

    def __repr__(self):
        return (f"{self.__class__.__name__}(in={self.in_channels}, out={self.out_channels}, "
                f"kernel_size={[self.kernel_size] * 3}, stride={[self.stride] * 3}, dilation=[1, 1, 1])")


class MinkowskiConvolution(MinkowskiConvolutionBase):
    is_transpose = False


class MinkowskiConvolutionTranspose(MinkowskiConvolutionBase):
    is_transpose = True
