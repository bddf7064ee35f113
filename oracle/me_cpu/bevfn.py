"""Oracle (test infrastructure): autograd wrapper of oracle.bev for the model."""
import numpy as np
import torch

from .. import bev as _bev


class _BevFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, coords, batch_size, bound, voxel_size, pool, policy):
        out, c = _bev.bev_forward(coords, feats.detach().numpy(), batch_size, bound, voxel_size, pool, policy)
        ctx.c = c
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad_out):
        return torch.from_numpy(_bev.bev_backward(grad_out.numpy(), ctx.c)), None, None, None, None, None, None


def sparse2super(x, bound=50.0, voxel_size=0.05, pool=(5, 3, 1), policy="last"):
    coords = x.C.numpy()
    batch_size = int(coords[:, 0].max()) + 1  # reference: batch_bottle_idx.max()+1 (minkunet_bev.py:193)
    return _BevFunction.apply(x.F, coords, batch_size, bound, voxel_size, pool, policy)
