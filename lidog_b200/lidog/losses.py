"""Device-resident DICE / SoftDICE losses.

Same arithmetic as the reference's `DICELoss` / `SoftDICELoss`
(utils/losses/losses.py:56-97, 129-187, soft targets :100-126) but without the
`.cpu()` round trips (`:72-73,148-149`) and without boolean-mask indexing, so no
host synchronisation: ignored rows are weighted by zero instead of being
removed, which gives the same sums.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

# 1 = on CUDA tensors the criteria run as two fused kernels each way (csrc/losses.cu: lg_dice_forward / _backward);
# 0 = the torch formulation below on every device (it is what the CPU oracle trainer evaluates, and the checker of
# the fused kernels in tests/test_golden_reference.py)
CONFIG = {"fused": int(os.environ.get("LIDOG_FUSED_LOSS", "1"))}


class _DiceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_label, soft, eps, is_kitti):
        from .. import cabi
        logits = logits.contiguous()
        target = target.contiguous()
        n, C = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        coef = torch.empty(2 * C, dtype=torch.float32, device=logits.device)
        cabi.check(cabi.lib().lg_dice_forward(logits.data_ptr(), target.data_ptr(), n, C, ignore_label, soft, eps, is_kitti,
                                              loss.data_ptr(), coef.data_ptr(), cabi.stream_of(logits)), "lg_dice_forward")
        ctx.save_for_backward(logits, target, coef)
        ctx.cfg = (ignore_label, soft, eps, is_kitti)
        return loss

    @staticmethod
    def backward(ctx, grad):
        from .. import cabi
        logits, target, coef = ctx.saved_tensors
        ignore_label, soft, eps, is_kitti = ctx.cfg
        n, C = logits.shape
        grad = grad.to(torch.float32).contiguous()
        dlogits = torch.empty_like(logits)
        cabi.check(cabi.lib().lg_dice_backward(logits.data_ptr(), target.data_ptr(), n, C, ignore_label, soft, eps, is_kitti,
                                               coef.data_ptr(), grad.data_ptr(), dlogits.data_ptr(),
                                               cabi.stream_of(logits)), "lg_dice_backward")
        return dlogits, None, None, None, None, None


def _fused_ok(output, target):
    return (CONFIG["fused"] and output.is_cuda and output.dtype == torch.float32 and output.dim() == 2
            and 2 <= output.shape[1] <= 32 and target.dtype == torch.int64)


_NO_IGNORE = -(1 << 30)  # ignore_label=None: no row is dropped


def _prep(output, target, ignore_label):
    C = output.shape[1]
    if ignore_label is None:
        valid = torch.ones_like(target, dtype=output.dtype)
    else:
        valid = (target != ignore_label).to(output.dtype)
    onehot = F.one_hot(target.clamp(0, C - 1), num_classes=C).to(output.dtype) * valid[:, None]
    prob = F.softmax(output, dim=-1) * valid[:, None]
    return C, valid, onehot, prob


def dice_loss(output, target, ignore_label=None):
    """DICELoss(powerize=False, use_tmask=False) -- LiDOG's BEV criterion."""
    if _fused_ok(output, target):
        return _DiceFunction.apply(output, target, _NO_IGNORE if ignore_label is None else int(ignore_label), 0, 0.0, 0)
    C, valid, onehot, prob = _prep(output, target, ignore_label)
    inter = (prob * onehot).sum(0)
    union = prob.sum(0) + onehot.sum(0) + 1e-12
    iou = (2 * inter / union).sum() / (C + 1e-12)
    return 1 - iou


def soft_dice_loss(output, target, ignore_label=None, eps=0.05, is_kitti=False):
    """SoftDICELoss(powerize=True, use_tmask=True, eps=0.05) -- LiDOG's 3D criterion."""
    if _fused_ok(output, target) and (not is_kitti or output.shape[1] > 6):
        return _DiceFunction.apply(output, target, _NO_IGNORE if ignore_label is None else int(ignore_label), 1,
                                   float(eps), 1 if is_kitti else 0)
    C, valid, onehot, prob = _prep(output, target, ignore_label)
    hi, lo = 1 - eps, eps / (C - 1)
    soft = onehot * hi + (valid[:, None] - onehot) * lo
    if is_kitti:  # get_kitti_soft: classes 1 and 6 share the mass
        amb = ((target == 6) | (target == 1)).to(output.dtype) * valid
        soft[:, 1] = soft[:, 1] * (1 - amb) + amb * hi / 2
        soft[:, 6] = soft[:, 6] * (1 - amb) + amb * hi / 2
    inter = (prob * soft).sum(0)
    union = prob.pow(2).sum(0) + soft.sum(0) + 1e-12
    tmask = (onehot.sum(0) > 0).to(output.dtype)
    iou = (tmask * 2 * inter / union).sum() / (tmask.sum() + 1e-12)
    return 1 - iou
