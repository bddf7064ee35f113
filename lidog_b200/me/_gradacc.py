"""Accumulate-in-place targets for gradients of tensors with several consumers.

A BasicBlock input x feeds conv1 AND the residual add (or conv1 and the downsample convolution): autograd sums the two
gradients with an extra elementwise pass over N x C floats (31 such adds, 0.9 ms per LiDOG training step,
profiles/r02_a_launch_summary.txt).  Here the first gradient that becomes available for x is OFFERED, keyed by the data
pointer of x; a tensor-core convolution whose input is that same x TAKES the offer, lets its dgrad epilogue add into
the offered tensor (lg_conv_layer_backward, accumulate_dx) and returns no gradient of its own.  The offered tensor is
also returned to autograd by its producer exactly as before, so an offer nobody takes changes nothing.  Offers are
dropped when a new batch starts."""
from __future__ import annotations

import os

import torch

CONFIG = {"enabled": int(os.environ.get("LIDOG_GRAD_ACCUM", "1"))}
_OFFERS = {}


def offer(x_ptr: int, grad: torch.Tensor) -> None:
    if CONFIG["enabled"] and grad is not None and grad.is_contiguous() and grad.dtype == torch.float32:
        _OFFERS[x_ptr] = grad


def take(x_ptr: int, shape) -> torch.Tensor | None:
    g = _OFFERS.pop(x_ptr, None)
    if g is None or tuple(g.shape) != tuple(shape):
        return None
    # The offer is only good while autograd still holds THAT tensor as (the start of) x's gradient.  If a third
    # consumer's gradient arrived in between, the engine summed into a new tensor and dropped the offered one: then
    # this table owns the last reference and adding into it would lose the contribution.
    if g._use_count() < 2:
        return None
    return g


def clear() -> None:
    _OFFERS.clear()
