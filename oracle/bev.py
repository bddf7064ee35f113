"""Oracle (test infrastructure): point-to-BEV projection, forward and backward.

Closed-form restatement of `MinkUNetBaseBEV.sparse2super` + `filter_bounds`
(reference utils/models/minkunet_bev.py:158-230), including its quirks
(SURVEY.md 8a-11, Appendix B.1/B.3):
  * pixel maths in three separately rounded float32 operations,
  * strict bounds, negative pixel_y wraps like a Python index,
  * the (H, W, C) buffer is re-viewed as (C, H, W) without a permute,
  * MaxPool2d(kernel, stride, pad) on the re-viewed tensor.
PINNED: tests/test_oracle_bev.py checks this file bit-exactly against vectors
produced by running the reference function itself on CPU
(tests/golden/make_bev_golden.py).

Duplicate-pixel policies
  'last': highest row index wins the pixel (what the reference's index_put_
          does single-threaded on CPU); backward hands the pixel's gradient to
          EVERY row that wrote it (index_put_ backward semantics).
  'max' : channel-wise max over the rows sharing the pixel (north_star's
          scatter-max); backward routes to the first row attaining the max.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def image_size(bound: float, voxel_size: float = 0.05) -> int:
    """max_height / max_width exactly as minkunet_bev.py:184-185 computes them."""
    return int(torch.tensor((bound - (-bound)) / voxel_size).int())


def pixel_indices(coords_xyz: np.ndarray, bound: float, voxel_size: float = 0.05):
    """(in_bounds[N] bool, px[N], py[N]) with the reference's float32 rounding
    (minkunet_bev.py:175,158-167,213-214)."""
    H = image_size(bound, voxel_size)
    vs = np.float32(voxel_size)
    lo = np.float32(-bound)
    hi = np.float32(bound)
    xyz = coords_xyz.astype(np.float32) * vs
    x, y = xyz[:, 0], xyz[:, 1]
    inb = (lo < x) & (x < hi) & (lo < y) & (y < hi)
    px = np.floor((x - lo) / vs).astype(np.int64)
    py = np.floor(np.float32(H) - (y - lo) / vs).astype(np.int64) - 1
    py = np.where(py < 0, py + H, py)  # python negative index wrap
    return inb, px, py


def bev_forward(coords: np.ndarray, feats: np.ndarray, batch_size: int, bound: float,
                voxel_size: float = 0.05, pool=(5, 3, 1), policy: str = "last"):
    """coords int [N,4] (b,x,y,z), feats f32 [N,C] -> (out [B,C,h,w] f32, ctx)."""
    assert policy in ("last", "max")
    coords = np.asarray(coords)
    feats = np.asarray(feats, dtype=np.float32)
    N, C = feats.shape
    H = W = image_size(bound, voxel_size)
    inb, px, py = pixel_indices(coords[:, 1:], bound, voxel_size)
    pix = py * W + px
    outs, ctxs = [], []
    for b in range(batch_size):
        rows = np.nonzero((coords[:, 0] == b) & inb)[0]
        p = pix[rows]
        dense = np.zeros((H * W, C), np.float32)
        if policy == "last":
            win = np.full(H * W, -1, np.int64)
            np.maximum.at(win, p, rows)
            wp = np.nonzero(win >= 0)[0]
            dense[wp] = feats[win[wp]]
            arg_row = None
        else:
            acc = np.full((H * W, C), -np.inf, np.float32)
            np.maximum.at(acc, p, feats[rows])
            touched = np.zeros(H * W, bool)
            touched[p] = True
            dense[touched] = acc[touched]
            # first row (ascending) attaining the max, per channel
            arg_row = np.full((H * W, C), -1, np.int64)
            for r, pr in zip(rows[::-1], p[::-1]):
                hit = feats[r] == dense[pr]
                arg_row[pr, hit] = r
            win = None
        scr = torch.from_numpy(dense.reshape(-1)).view(1, C, H, W)  # raw re-view, no permute
        out, idx = F.max_pool2d(scr, pool[0], pool[1], pool[2], return_indices=True)
        outs.append(out.numpy())
        ctxs.append(dict(rows=rows, pix=p, win=win, arg_row=arg_row, idx=idx.numpy()))
    ctx = dict(per_sample=ctxs, N=N, C=C, H=H, W=W, policy=policy)
    return np.concatenate(outs, axis=0), ctx


def bev_backward(grad_out: np.ndarray, ctx) -> np.ndarray:
    """d(feats) [N,C] for a grad on the [B,C,h,w] output."""
    N, C, H, W = ctx["N"], ctx["C"], ctx["H"], ctx["W"]
    policy = ctx["policy"]
    dF = np.zeros((N, C), np.float32)
    for b, s in enumerate(ctx["per_sample"]):
        g = np.asarray(grad_out[b], np.float32).reshape(C, -1)
        idx = s["idx"].reshape(C, -1)
        gscr = np.zeros(C * H * W, np.float32)
        flat_idx = (np.arange(C)[:, None] * (H * W) + idx).reshape(-1)
        np.add.at(gscr, flat_idx, g.reshape(-1))
        gpix = gscr.reshape(H * W, C)
        rows, p = s["rows"], s["pix"]
        if policy == "last":
            dF[rows] = gpix[p]
        else:
            ar = s["arg_row"][p]
            dF[rows] = np.where(ar == rows[:, None], gpix[p], 0.0).astype(np.float32)
    return dF


def bev_forward_sparse_reference(coords, feats, batch_size, bound, voxel_size=0.05, pool=(5, 3, 1)):
    """Line-by-line functional twin of the reference's torch code (used only to
    cross-check `bev_forward(policy='last')` when /root/reference is absent)."""
    torch.set_num_threads(1)
    C = feats.shape[1]
    H = image_size(bound, voxel_size)
    c = torch.as_tensor(coords)
    f = torch.as_tensor(feats)
    xyz = c[:, 1:] * voxel_size
    outs = []
    for b in range(batch_size):
        m = c[:, 0] == b
        fb, xb = f[m], xyz[m]
        inb = (-bound < xb[:, 0]) & (xb[:, 0] < bound) & (-bound < xb[:, 1]) & (xb[:, 1] < bound)
        fb, xb = fb[inb], xb[inb]
        img = torch.zeros((H, H, C))
        px = torch.floor((xb[:, 0] - (-bound)) / voxel_size).long()
        py = torch.floor(torch.tensor(H).int() - (xb[:, 1] - (-bound)) / voxel_size).long() - 1
        img[py, px] = fb
        outs.append(F.max_pool2d(img.view(1, -1, H, H), pool[0], pool[1], pool[2]))
    return torch.cat(outs, 0).numpy()
