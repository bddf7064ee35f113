"""Oracle (test infrastructure): CPU SparseTensor / CoordinateManager / layers.

A minimal MinkowskiEngine-0.5.4-shaped API on top of oracle.voxel / oracle.conv
(numpy + torch CPU).  Conventions: SURVEY.md Appendix C.  Call sites this must
satisfy: trainer_lighting_2d.py:151 (SparseTensor), minkunet_bev.py:44-157
(layer ctors, `.kernel`, `.bn`), :172,211 (`.C`, `.F`, `.device`), :337-370
(`ME.cat`), train_lidog.py:228 (convert_sync_batchnorm).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from .. import voxel as vox
from ..conv import SparseConvFunction


class CoordinateManager:
    """Owns the coordinate sets per tensor stride and the kernel-map cache."""

    def __init__(self, coords_ts1: np.ndarray):
        self.coords = {1: np.ascontiguousarray(coords_ts1, dtype=np.int32)}
        self.parent = {}
        self.maps = {}

    def get_coords(self, ts: int) -> np.ndarray:
        if ts not in self.coords:
            half = ts // 2
            assert half * 2 == ts and half >= 1
            c, inv = vox.stride_coords(self.get_coords(half), ts)
            self.coords[ts], self.parent[ts] = c, inv
        return self.coords[ts]

    def kernel_map(self, ts_in: int, ts_out: int, kernel_size: int, transposed: bool):
        key = (ts_in, ts_out, kernel_size, transposed)
        if key not in self.maps:
            if kernel_size == 1 and ts_in == ts_out:
                n = self.get_coords(ts_in).shape[0]
                r = np.arange(n, dtype=np.int64)
                self.maps[key] = [(r, r)]
            elif transposed:
                # in = coarse (ts_in), out = fine (ts_out)
                self.maps[key] = vox.transposed_kernel_map(self.get_coords(ts_out), self.get_coords(ts_in),
                                                           kernel_size, ts_out)
            else:
                self.maps[key] = vox.kernel_map(self.get_coords(ts_in), self.get_coords(ts_out), kernel_size, ts_in)
            self.maps[key] = [(torch.from_numpy(i), torch.from_numpy(o)) for i, o in self.maps[key]]
        return self.maps[key]


class SparseTensor:
    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_manager=None, device=None):
        if coordinate_manager is None:
            assert coordinates is not None
            c = coordinates.detach().cpu().numpy() if isinstance(coordinates, torch.Tensor) else np.asarray(coordinates)
            c = c.astype(np.int32)
            umap, _ = vox.unique_first_occurrence(c)
            if umap.shape[0] != c.shape[0]:  # duplicates: keep first occurrences
                c = c[umap]
                features = features[torch.from_numpy(umap)]
            coordinate_manager = CoordinateManager(c)
        self._F = features
        self.coordinate_manager = coordinate_manager
        self.tensor_stride = int(tensor_stride)

    @property
    def F(self):
        return self._F

    @property
    def C(self):
        return torch.from_numpy(self.coordinate_manager.get_coords(self.tensor_stride))

    @property
    def device(self):
        return self._F.device

    @property
    def shape(self):
        return self._F.shape

    def _like(self, feats):
        return SparseTensor(feats, tensor_stride=self.tensor_stride, coordinate_manager=self.coordinate_manager)

    def __add__(self, other):
        return self._like(self._F + (other._F if isinstance(other, SparseTensor) else other))

    def __iadd__(self, other):
        self._F = self._F + (other._F if isinstance(other, SparseTensor) else other)
        return self


def cat(*tensors):
    t0 = tensors[0]
    for t in tensors[1:]:
        assert t.coordinate_manager is t0.coordinate_manager and t.tensor_stride == t0.tensor_stride
    return t0._like(torch.cat([t.F for t in tensors], dim=1))


class _ConvBase(nn.Module):
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, dimension=3):
        super().__init__()
        assert dimension == 3 and dilation == 1
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dimension = int(kernel_size), int(stride), dimension
        K = self.kernel_size ** 3
        shape = (in_channels, out_channels) if K == 1 else (K, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        K = self.kernel_size ** 3
        fan = (self.out_channels if self.transposed else self.in_channels) * K
        bound = 1.0 / math.sqrt(fan)
        with torch.no_grad():
            self.kernel.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, x: SparseTensor) -> SparseTensor:
        cm = x.coordinate_manager
        ts_in = x.tensor_stride
        ts_out = ts_in // self.stride if self.transposed else ts_in * self.stride
        n_out = cm.get_coords(ts_out).shape[0]
        maps = cm.kernel_map(ts_in, ts_out, self.kernel_size, self.transposed)
        y = SparseConvFunction.apply(x.F, self.kernel, maps, n_out)
        if self.bias is not None:
            y = y + self.bias
        return SparseTensor(y, tensor_stride=ts_out, coordinate_manager=cm)


class MinkowskiConvolution(_ConvBase):
    transposed = False


class MinkowskiConvolutionTranspose(_ConvBase):
    transposed = True


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return x._like(self.bn(x.F))


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 process_group=None):
        nn.Module.__init__(self)
        self.bn = nn.SyncBatchNorm(num_features, eps=eps, momentum=momentum, affine=affine,
                                   track_running_stats=track_running_stats, process_group=process_group)

    @classmethod
    def convert_sync_batchnorm(cls, module, process_group=None):
        out = module
        if isinstance(module, MinkowskiBatchNorm) and not isinstance(module, MinkowskiSyncBatchNorm):
            out = cls(module.bn.num_features, module.bn.eps, module.bn.momentum, module.bn.affine,
                      module.bn.track_running_stats, process_group)
            if module.bn.affine:
                with torch.no_grad():
                    out.bn.weight, out.bn.bias = module.bn.weight, module.bn.bias
            out.bn.running_mean, out.bn.running_var = module.bn.running_mean, module.bn.running_var
            out.bn.num_batches_tracked = module.bn.num_batches_tracked
            return out
        for name, child in module.named_children():
            out.add_module(name, cls.convert_sync_batchnorm(child, process_group))
        return out


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x: SparseTensor) -> SparseTensor:
        return x._like(torch.relu(x.F))


class MinkowskiDropout(nn.Module):
    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.drop = nn.Dropout(p)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return x._like(self.drop(x.F))
