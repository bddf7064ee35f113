"""MinkowskiBatchNorm / MinkowskiSyncBatchNorm / ReLU / Dropout.

ME implements these as torch modules applied to the N x C feature matrix
(`.bn` attribute: utils/models/minkunet_bev.py:407-408; recursive
`convert_sync_batchnorm`: train_lidog.py:228).  The statistics, affine and
activation stay torch/cuDNN/NCCL library calls in this round (SURVEY.md 8f-1
lists their fusion as the next row).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .sparse_tensor import SparseTensor


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, input: SparseTensor) -> SparseTensor:
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
        return input._like(torch.relu_(input.F) if self.inplace else torch.relu(input.F))


class MinkowskiDropout(nn.Module):
    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.module = nn.Dropout(p, inplace)

    def forward(self, input: SparseTensor) -> SparseTensor:
        return input._like(self.module(input.F))
