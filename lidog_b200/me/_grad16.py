"""Side channel between the fused BN backward and the convolution backward.

The BN backward (me/norm.py) produces the gradient of a convolution output twice: fp32 for autograd
and as the scaled 16-bit copy the tensor-core dgrad / wgrad kernels read.  autograd only carries the
fp32 tensor, so the 16-bit copy waits here, keyed by the fp32 tensor's data pointer; the convolution's
backward pops its entry instead of running absmax + cast again.  An entry keeps its fp32 tensor alive
(the pointer cannot be recycled while listed) and the table is emptied when a new batch starts.

`fp32_valid=False`: the BN backward did not write the fp32 tensor at all (its only consumer is a tensor-core
convolution, which reads the 16-bit copy: 4 of 14-18 bytes per element of that pass).  The convolution's backward
refuses such a gradient when the 16-bit copy cannot be used (`require_fp32`), and the BN layer watches the gradient of
its input with a tensor hook (me/norm.py) so that a second consumer of the same tensor fails loudly instead of
summing uninitialised memory.
"""
from __future__ import annotations

import torch

_TABLE = {}
_NO_FP32 = set()  # data pointers of published gradients whose fp32 tensor was never written


def publish_grad16(grad: torch.Tensor, grad16: torch.Tensor, scale: torch.Tensor, fmt: int, fp32_valid: bool = True) -> None:
    _TABLE[grad.data_ptr()] = (grad, grad16, scale, fmt)
    if not fp32_valid:
        _NO_FP32.add(grad.data_ptr())


def take_grad16(grad: torch.Tensor, fmt: int):
    """(grad16, scale[4]) for exactly this gradient tensor, or None."""
    hit = _TABLE.pop(grad.data_ptr(), None)
    if hit is None:
        return None
    g, g16, scale, f = hit
    if g.shape != grad.shape or g._version != grad._version or g.stride() != grad.stride() or f != fmt:
        return None
    _NO_FP32.discard(grad.data_ptr())
    return g16, scale


def require_fp32(grad: torch.Tensor, who: str) -> None:
    """Called by a consumer that is about to READ the fp32 values of `grad`."""
    if grad.data_ptr() in _NO_FP32:
        raise RuntimeError(f"{who}: this gradient was published in 16 bits only (fused BN backward, "
                           "LIDOG_BN_SKIP_DX32=1) and its fp32 values were never written; set LIDOG_BN_SKIP_DX32=0")


def clear() -> None:
    _TABLE.clear()
    _NO_FP32.clear()
