"""MinkowskiBatchNorm / MinkowskiSyncBatchNorm / ReLU / Dropout.

ME implements these as torch modules applied to the N x C feature matrix
(`.bn` attribute: utils/models/minkunet_bev.py:407-408; recursive
`convert_sync_batchnorm`: train_lidog.py:228).  The parameters and buffers still live in
`self.bn` (a torch BatchNorm1d / SyncBatchNorm, so state dicts stay ME-compatible), but in
training mode the arithmetic runs in liblidog_b200 (csrc/bn.cu), fused with what the
reference chains around it (utils/models/minkunet_bev.py:308-368, ME's BasicBlock):

    conv -> BN -> ReLU                         one statistics pass + one apply pass
    conv -> BN -> (+= residual) -> ReLU        the residual (a plain tensor or a second,
                                               not yet applied BN: the block's downsample)
                                               joins the same apply pass

The module API is unchanged, so the fusion is done lazily: `MinkowskiBatchNorm.forward`
returns a `DeferredBN` sparse tensor that only remembers its inputs; `+=` attaches the
residual, `MinkowskiReLU` (or any read of `.F`) runs the fused kernels.  The apply pass
also writes the 16-bit operand copy the next tensor-core convolution gathers, and the
backward pass hands the scaled 16-bit gradient straight to the convolution's dgrad / wgrad.
SyncBN exchanges the (2C+1) sums of the statistics pass (forward) and the 2C sums of the
backward pass with one all-reduce each.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import cabi
from .sparse_tensor import SparseTensor

CONFIG = {"fused": int(os.environ.get("LIDOG_FUSED_BN", "1"))}

from ._grad16 import publish_grad16


def _fmt16():
    from . import conv as meconv
    return meconv._fmt()


def _dtype16(fmt):
    return torch.float16 if fmt == cabi.FMT_FP16 else torch.bfloat16


def _group(bn):
    """(process group, world size) for a SyncBatchNorm in a live process group, else (None, 1)."""
    if isinstance(bn, nn.SyncBatchNorm) and torch.distributed.is_available() and torch.distributed.is_initialized():
        pg = bn.process_group if bn.process_group is not None else torch.distributed.group.WORLD
        ws = torch.distributed.get_world_size(pg)
        if ws > 1:
            return pg, ws
    return None, 1


def _all_sum(v: torch.Tensor, pg) -> torch.Tensor:
    """Sum of a float64 vector over the ranks of `pg`: the one-kernel peer-memory exchange (me/peer.py) when the
    node offers symmetric memory, an NCCL all-reduce otherwise.  Returns a new tensor."""
    from . import peer
    ex = peer.get(pg)
    out = torch.empty_like(v)
    if ex is not None:
        return ex.sum(v, out)
    out.copy_(v)
    torch.distributed.all_reduce(out, group=pg)
    return out


def _bn_statistics(x: torch.Tensor, bn, ws, ws_bytes):
    """Training-mode statistics of one BN: returns (stats [4C] = mean, invstd, scale, shift; global count; group)."""
    L = cabi.lib()
    n, C = x.shape
    dev = x.device
    sums = torch.empty(2 * C + 1, dtype=torch.float64, device=dev)
    cabi.check(L.lg_bn_stats(cabi.ptr(x), n, C, cabi.ptr(sums), cabi.ptr(ws), ws_bytes, cabi.stream()), "lg_bn_stats")
    pg, world = _group(bn)
    count = float(n)  # a host number without SyncBN, a device scalar (never read back) with it
    if pg is not None:  # SyncBN: one exchange of [sum x, sum x^2, n]
        sums = _all_sum(sums, pg)
        count = sums[2 * C:]
    count_host, count_dev = (count, None) if isinstance(count, float) else (0.0, count)
    stats = torch.empty(4 * C, dtype=torch.float32, device=dev)
    track = bn.track_running_stats and bn.running_mean is not None
    momentum = 0.1 if bn.momentum is None else float(bn.momentum)
    cabi.check(L.lg_bn_finalize(cabi.ptr(sums), count_host, cabi.ptr(count_dev), C,
                                cabi.ptr(bn.weight.detach() if bn.affine else None),
                                cabi.ptr(bn.bias.detach() if bn.affine else None), float(bn.eps), momentum,
                                cabi.ptr(bn.running_mean if track else None),
                                cabi.ptr(bn.running_var if track else None),
                                cabi.ptr(bn.num_batches_tracked if track else None), cabi.ptr(stats), cabi.stream()),
               "lg_bn_finalize")
    return stats, count, pg


class FusedBNFunction(torch.autograd.Function):
    """y = act(BN_a(x) [+ BN_b(x2)] [+ res]); also returns nothing else -- the 16-bit copy of y travels in `box`."""

    @staticmethod
    def forward(ctx, x, w, b, x2, w2, b2, res, bn_a, bn_b, relu, box):
        L = cabi.lib()
        x = x.contiguous()
        n, C = x.shape
        dev = x.device
        ws_bytes = L.lg_bn_workspace(n, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        st_a, count_a, pg_a = _bn_statistics(x, bn_a, ws, ws_bytes)
        st_b = count_b = pg_b = None
        if x2 is not None:
            x2 = x2.contiguous()
            st_b, count_b, pg_b = _bn_statistics(x2, bn_b, ws, ws_bytes)
        if res is not None:
            res = res.contiguous()
        fmt = _fmt16()
        y = torch.empty_like(x)
        y16 = torch.empty(x.shape, dtype=_dtype16(fmt), device=dev) if fmt is not None else None
        cabi.check(L.lg_bn_apply(cabi.ptr(x), cabi.ptr(st_a), cabi.ptr(x2), cabi.ptr(st_b), cabi.ptr(res), int(relu), n, C,
                                 cabi.ptr(y), cabi.ptr(y16), fmt if fmt is not None else 0, cabi.stream()), "lg_bn_apply")
        box["y16"], box["fmt"] = y16, fmt
        ctx.save_for_backward(x, x2, y if relu else None, st_a, st_b,
                              w.detach() if w is not None else None, w2.detach() if w2 is not None else None)
        ctx.meta = (relu, count_a, pg_a, count_b, pg_b, res is not None, fmt)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = cabi.lib()
        x, x2, y, st_a, st_b, w, w2 = ctx.saved_tensors
        relu, count_a, pg_a, count_b, pg_b, has_res, fmt = ctx.meta
        dy = dy.contiguous()
        n, C = x.shape
        dev = x.device
        ws_bytes = L.lg_bn_workspace(n, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        sums = torch.empty(3 * C, dtype=torch.float64, device=dev)
        maxes = torch.empty(3 * C, dtype=torch.float32, device=dev)
        cabi.check(L.lg_bn_bwd_stats(cabi.ptr(dy), cabi.ptr(y), cabi.ptr(x), cabi.ptr(st_a), cabi.ptr(x2),
                                     cabi.ptr(st_b), int(relu), n, C, cabi.ptr(sums), cabi.ptr(maxes), cabi.ptr(ws),
                                     ws_bytes, cabi.stream()), "lg_bn_bwd_stats")
        sums_g = sums
        pg = pg_a if pg_a is not None else pg_b
        if pg is not None:  # SyncBN: dx needs the sums over every rank; dgamma / dbeta stay local (DDP reduces them)
            sums_g = _all_sum(sums, pg)  # (the fp16 bound only has to hold on this rank: maxes stay local)
        use16 = fmt is not None
        d16 = _dtype16(fmt) if use16 else None

        def branch(k, st, gamma, count, want_param_grads):
            coef = torch.empty(3 * C, dtype=torch.float32, device=dev)
            scale = torch.empty(4, dtype=torch.float32, device=dev)
            dgamma = torch.empty(C, dtype=torch.float32, device=dev) if want_param_grads else None
            dbeta = torch.empty(C, dtype=torch.float32, device=dev) if want_param_grads else None
            count_host, count_dev = (count, None) if isinstance(count, float) else (0.0, count)
            cabi.check(L.lg_bn_bwd_finalize(cabi.ptr(sums), cabi.ptr(sums_g), cabi.ptr(maxes), count_host,
                                            cabi.ptr(count_dev), C, k,
                                            cabi.ptr(gamma), cabi.ptr(st), cabi.ptr(coef), cabi.ptr(dgamma),
                                            cabi.ptr(dbeta), cabi.ptr(scale), cabi.stream()), "lg_bn_bwd_finalize")
            return coef, scale, dgamma, dbeta

        coef_a, scale_a, dw, db = branch(0, st_a, w, count_a, w is not None)
        dx = torch.empty_like(x)
        dx16 = torch.empty(x.shape, dtype=d16, device=dev) if use16 else None
        coef_b = scale_b = dw2 = db2 = dx2 = dx2_16 = None
        if x2 is not None:
            coef_b, scale_b, dw2, db2 = branch(1, st_b, w2, count_b, w2 is not None)
            dx2 = torch.empty_like(x2)
            dx2_16 = torch.empty(x2.shape, dtype=d16, device=dev) if use16 else None
        dres = dres16 = scale_r = None
        if has_res and ctx.needs_input_grad[6]:
            dres = torch.empty_like(x)
            if use16:
                dres16 = torch.empty(x.shape, dtype=d16, device=dev)
                scale_r = torch.empty(4, dtype=torch.float32, device=dev)
                cabi.check(L.lg_bn_bwd_gscale(cabi.ptr(maxes), C, cabi.ptr(scale_r), cabi.stream()), "lg_bn_bwd_gscale")
        cabi.check(L.lg_bn_bwd_apply(cabi.ptr(dy), cabi.ptr(y), cabi.ptr(x), cabi.ptr(st_a), cabi.ptr(coef_a),
                                     cabi.ptr(x2), cabi.ptr(st_b), cabi.ptr(coef_b), int(relu), n, C, cabi.ptr(dx),
                                     cabi.ptr(dx16), cabi.ptr(scale_a), cabi.ptr(dx2), cabi.ptr(dx2_16),
                                     cabi.ptr(scale_b), cabi.ptr(dres), cabi.ptr(dres16), cabi.ptr(scale_r),
                                     fmt if use16 else 0, cabi.stream()), "lg_bn_bwd_apply")
        if use16:
            publish_grad16(dx, dx16, scale_a, fmt)
            if dx2 is not None:
                publish_grad16(dx2, dx2_16, scale_b, fmt)
            if dres is not None:
                publish_grad16(dres, dres16, scale_r, fmt)
        return dx, dw, db, dx2, dw2, db2, dres, None, None, None, None


def _fusable(bn, feats: torch.Tensor) -> bool:
    return (CONFIG["fused"] and bn.training and feats.is_cuda and feats.dtype == torch.float32 and feats.dim() == 2
            and feats.shape[0] > 1 and feats.shape[1] % 4 == 0 and feats.shape[1] <= 1024 and bn.affine
            and (bn.momentum is not None or not bn.track_running_stats))


class DeferredBN(SparseTensor):
    """BN output that has not been computed yet (see the module docstring).  Behaves as a SparseTensor; the first
    read of `.F` evaluates BN(src) [+ residual] without activation."""

    def __init__(self, src: SparseTensor, bn):
        self.coordinate_manager = src.coordinate_manager
        self._ts = src._ts
        self._f16_cache = None
        self._src, self._bn = src, bn
        self._residual = None
        self._value = None

    @property
    def _F(self):
        if self._value is None:
            self._evaluate(False)
        return self._value

    @_F.setter
    def _F(self, v):
        self._value = v

    def _evaluate(self, relu: bool):
        bn = self._bn
        x2 = w2 = b2 = res = bn_b = None
        r = self._residual
        if isinstance(r, DeferredBN) and r._value is None and r._residual is None:
            x2, w2, b2, bn_b = r._src.F, r._bn.weight, r._bn.bias, r._bn
        elif r is not None:
            res = r.F if isinstance(r, SparseTensor) else r
        box = {}
        self._value = FusedBNFunction.apply(self._src.F, bn.weight, bn.bias, x2, w2, b2, res, bn, bn_b, bool(relu), box)
        if box.get("y16") is not None:  # the next convolution's operand: no separate cast pass
            f = self._value
            self._f16_cache = ((box["fmt"], f._version, f.data_ptr()), box["y16"])
        self._src = self._residual = None
        return self

    def __iadd__(self, other):
        if self._value is None and self._residual is None and isinstance(other, SparseTensor):
            self._same_map(other)
            self._residual = other
            return self
        return SparseTensor.__iadd__(self, other)


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, input: SparseTensor) -> SparseTensor:
        if _fusable(self.bn, input.F):
            return DeferredBN(input, self.bn)
        return input._like(self.bn(input.F))

    def __repr__(self):
        b = self.bn
        return f"{self.__class__.__name__}({b.num_features}, eps={b.eps}, momentum={b.momentum}, affine={b.affine})"


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 process_group=None):
        nn.Module.__init__(self)
        self.bn = nn.SyncBatchNorm(num_features, eps=eps, momentum=momentum, affine=affine,
                                   track_running_stats=track_running_stats, process_group=process_group)

    @classmethod
    def convert_sync_batchnorm(cls, module, process_group=None):
        """Recursively swap MinkowskiBatchNorm for MinkowskiSyncBatchNorm, keeping parameters and buffers."""
        out = module
        if isinstance(module, MinkowskiBatchNorm) and not isinstance(module, MinkowskiSyncBatchNorm):
            b = module.bn
            out = cls(b.num_features, b.eps, b.momentum, b.affine, b.track_running_stats, process_group)
            if b.affine:
                with torch.no_grad():
                    out.bn.weight, out.bn.bias = b.weight, b.bias
            out.bn.running_mean, out.bn.running_var = b.running_mean, b.running_var
            out.bn.num_batches_tracked = b.num_batches_tracked
            if hasattr(b, "qconfig"):
                out.bn.qconfig = b.qconfig
            return out
        for name, child in module.named_children():
            out.add_module(name, cls.convert_sync_batchnorm(child, process_group))
        return out


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, input: SparseTensor) -> SparseTensor:
        if isinstance(input, DeferredBN) and input._value is None:
            return input._evaluate(True)
        return input._like(torch.relu_(input.F) if self.inplace else torch.relu(input.F))


class MinkowskiDropout(nn.Module):
    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.module = nn.Dropout(p, inplace)

    def forward(self, input: SparseTensor) -> SparseTensor:
        return input._like(self.module(input.F))
