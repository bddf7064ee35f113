"""Side channel between the fused BN backward and the convolution backward.

The BN backward (me/norm.py) produces the gradient of a convolution output twice: fp32 for autograd
and as the scaled 16-bit copy the tensor-core dgrad / wgrad kernels read.  autograd only carries the
fp32 tensor, so the 16-bit copy waits here, keyed by the fp32 tensor's data pointer; the convolution's
backward pops its entry instead of running absmax + cast again.  An entry keeps its fp32 tensor alive
(the pointer cannot be recycled while listed) and the table is emptied when a new batch starts.

`fp32_valid=False`: the BN backward did not write the fp32 tensor at all (its only consumer is a tensor-core
convolution, which reads the 16-bit copy: 4 of 14-18 bytes per element of that pass).  The convolution's backward
refuses such a gradient when the 16-bit copy cannot be used, and recognises the sum autograd builds when the
convolution output has a second consumer (`require_fp32`): both fail loudly instead of reading uninitialised memory.
"""
from __future__ import annotations

import torch

_TABLE = {}
_NO_FP32 = set()  # data pointers of published gradients whose fp32 tensor was never written


def publish_grad16(grad: torch.Tensor, grad16: torch.Tensor, scale: torch.Tensor, fmt: int, fp32_valid: bool = True) -> None:
    _TABLE[grad.data_ptr()] = (grad, grad16, scale, fmt, grad._version)  # version at publication time
    if not fp32_valid:
        _NO_FP32.add(grad.data_ptr())


def take_grad16(grad: torch.Tensor, fmt: int):
    """(grad16, scale[4]) for exactly this gradient tensor, or None."""
    hit = _TABLE.get(grad.data_ptr())
    if hit is None:
        return None
    g, g16, scale, f, version = hit
    if g.shape != grad.shape or version != grad._version or g.stride() != grad.stride() or f != fmt:
        if grad.data_ptr() not in _NO_FP32:
            del _TABLE[grad.data_ptr()]  # fp32 values exist: the caller casts them itself
        return None
    del _TABLE[grad.data_ptr()]
    _NO_FP32.discard(grad.data_ptr())
    return g16, scale


def require_fp32(grad: torch.Tensor, who: str) -> None:
    """Called by a consumer that is about to READ the fp32 values of `grad`.  Refuses (a) a gradient that was
    published in 16 bits only, and (b) any gradient of the same shape as a 16-bit-only one that is still waiting for
    its convolution: that is what a second consumer of the convolution output looks like from here -- autograd summed
    the unwritten tensor with the other consumer's gradient (in place: the version check of `take_grad16` turned the
    hit into a miss; out of place: a new tensor arrived and the published one was never taken)."""
    bad = grad.data_ptr() in _NO_FP32
    if not bad:
        for ptr in _NO_FP32:
            entry = _TABLE.get(ptr)
            if entry is not None and entry[0].shape == grad.shape:
                bad = True
                break
    if bad:
        raise RuntimeError(f"{who}: a gradient of this convolution output was published in 16 bits only (fused BN "
                           "backward) and its fp32 values were never written, but the fp32 values are needed here -- "
                           "the output has a second consumer or the 16-bit copy is unusable; set LIDOG_BN_SKIP_DX32=0")


def clear() -> None:
    _TABLE.clear()
    _NO_FP32.clear()
