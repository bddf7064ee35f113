"""Device-resident DICE / SoftDICE losses.

Same arithmetic as the reference's `DICELoss` / `SoftDICELoss`
(utils/losses/losses.py:56-97, 129-187, soft targets :100-126) but without the
`.cpu()` round trips (`:72-73,148-149`) and without boolean-mask indexing, so no
host synchronisation: ignored rows are weighted by zero instead of being
removed, which gives the same sums.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


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
    C, valid, onehot, prob = _prep(output, target, ignore_label)
    inter = (prob * onehot).sum(0)
    union = prob.sum(0) + onehot.sum(0) + 1e-12
    iou = (2 * inter / union).sum() / (C + 1e-12)
    return 1 - iou


def soft_dice_loss(output, target, ignore_label=None, eps=0.05, is_kitti=False):
    """SoftDICELoss(powerize=True, use_tmask=True, eps=0.05) -- LiDOG's 3D criterion."""
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
