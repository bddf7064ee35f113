"""Side channel between the fused BN backward and the convolution backward.

The BN backward (me/norm.py) produces the gradient of a convolution output twice: fp32 for autograd
and as the scaled 16-bit copy the tensor-core dgrad / wgrad kernels read.  autograd only carries the
fp32 tensor, so the 16-bit copy waits here, keyed by the fp32 tensor's data pointer; the convolution's
backward pops its entry instead of running absmax + cast again.  An entry keeps its fp32 tensor alive
(the pointer cannot be recycled while listed) and the table is emptied when a new batch starts.
"""
from __future__ import annotations

import torch

_TABLE = {}


def publish_grad16(grad: torch.Tensor, grad16: torch.Tensor, scale: torch.Tensor, fmt: int) -> None:
    _TABLE[grad.data_ptr()] = (grad, grad16, scale, fmt)


def take_grad16(grad: torch.Tensor, fmt: int):
    """(grad16, scale[4]) for exactly this gradient tensor, or None."""
    hit = _TABLE.pop(grad.data_ptr(), None)
    if hit is None:
        return None
    g, g16, scale, f = hit
    if g.shape != grad.shape or g._version != grad._version or g.stride() != grad.stride() or f != fmt:
        return None
    return g16, scale


def clear() -> None:
    _TABLE.clear()
