"""Oracle (test infrastructure): sparse convolution forward / dgrad / wgrad.

Restates the structure of MinkowskiEngine's CPU convolution backend
(per kernel offset: gather rows -> dense matmul -> scatter-add), as summarised
in SURVEY.md section 8a rows a-6..a-9.  ME 0.5.4 is absent => parity unpinned;
tests additionally check this file against torch.nn.functional.conv3d on a
densified grid (an ME-independent ground truth).

All functions work in the dtype of their inputs (float32 = the reference's
precision, float64 = ground truth for tolerance checks).
"""
from __future__ import annotations

import torch


def _as_index(a):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(a)


def conv_forward(x: torch.Tensor, w: torch.Tensor, maps, n_out: int) -> torch.Tensor:
    """Y[o] = sum_k sum_{(i,o) in M_k} X[i] @ W[k];  w is (K, Cin, Cout)."""
    if w.dim() == 2:
        w = w.unsqueeze(0)
    y = torch.zeros((n_out, w.shape[2]), dtype=x.dtype)
    for k, (in_rows, out_rows) in enumerate(maps):
        in_rows, out_rows = _as_index(in_rows), _as_index(out_rows)
        if in_rows.numel() == 0:
            continue
        y.index_add_(0, out_rows, x.index_select(0, in_rows) @ w[k])
    return y


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, maps, n_in: int) -> torch.Tensor:
    """dX[i] += dY[o] @ W[k]^T over the pairs of each offset."""
    if w.dim() == 2:
        w = w.unsqueeze(0)
    dx = torch.zeros((n_in, w.shape[1]), dtype=dy.dtype)
    for k, (in_rows, out_rows) in enumerate(maps):
        in_rows, out_rows = _as_index(in_rows), _as_index(out_rows)
        if in_rows.numel() == 0:
            continue
        dx.index_add_(0, in_rows, dy.index_select(0, out_rows) @ w[k].t())
    return dx


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, maps, K: int) -> torch.Tensor:
    """dW[k] = X[in_k]^T @ dY[out_k]."""
    dw = torch.zeros((K, x.shape[1], dy.shape[1]), dtype=x.dtype)
    for k, (in_rows, out_rows) in enumerate(maps):
        in_rows, out_rows = _as_index(in_rows), _as_index(out_rows)
        if in_rows.numel() == 0:
            continue
        dw[k] = x.index_select(0, in_rows).t() @ dy.index_select(0, out_rows)
    return dw


class SparseConvFunction(torch.autograd.Function):
    """Autograd wrapper used by the oracle's MinkowskiEngine stand-in."""

    @staticmethod
    def forward(ctx, x, w, maps, n_out):
        ctx.maps = maps
        ctx.n_in = x.shape[0]
        ctx.w_dim = w.dim()
        ctx.save_for_backward(x, w)
        return conv_forward(x, w, maps, n_out)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = conv_dgrad(dy, w, ctx.maps, ctx.n_in) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            K = 1 if w.dim() == 2 else w.shape[0]
            dw = conv_wgrad(x, dy, ctx.maps, K)
            if ctx.w_dim == 2:
                dw = dw[0]
        return dx, dw, None, None
