"""ME.utils surface used by LiDOG: sparse_quantize, SparseCollation, batched_coordinates,
kaiming_normal_.  Voxelisation runs on the GPU (csrc/coords.cu); numpy / CPU-tensor inputs are
copied to the device and the results copied back in the input's container type, so the dataset
call sites (utils/datasets/semantickitti_bev.py:232-238, synth4d_bev.py:274-280,
nuscenes_bev.py:244-250, mix3D.py:67-72) work unchanged.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from ... import cabi
from ..coords import build_levels, coords_unique


def _device(device=None):
    if device is not None and str(device) != "cpu":
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("lidog_b200 has no CPU path: sparse_quantize needs a CUDA device")
    return torch.device("cuda", torch.cuda.current_device())


def quantize_points(points: torch.Tensor, quantization_size, batch_of_row: torch.Tensor | None = None) -> torch.Tensor:
    """[N,3] device points -> int32 [N,4] (batch, floor(p / size)) on the device, in the precision numpy would use:
    float64 points (the reference's augmented training clouds) divide in float64, everything else in float32."""
    assert points.is_cuda and points.dim() == 2 and points.shape[1] == 3
    f64 = points.dtype == torch.float64
    pts = points.contiguous() if f64 else points.to(torch.float32).contiguous()
    cast = float if f64 else (lambda v: float(np.float32(v)))
    if isinstance(quantization_size, (int, float)):
        sx = sy = sz = cast(quantization_size)
    else:
        sx, sy, sz = (cast(v) for v in quantization_size)
    out = torch.empty((pts.shape[0], 4), dtype=torch.int32, device=pts.device)
    if batch_of_row is not None:
        batch_of_row = batch_of_row.to(device=pts.device, dtype=torch.int32).contiguous()
    L = cabi.lib()
    fn, name = (L.lg_quantize_points_f64, "lg_quantize_points_f64") if f64 else (L.lg_quantize_points, "lg_quantize_points")
    cabi.check(fn(cabi.ptr(pts), cabi.ptr(batch_of_row), pts.shape[0], sx, sy, sz, cabi.ptr(out), cabi.stream_of(pts)), name)
    return out


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device=None):
    """GPU voxelisation with ME's contract: unique voxels in first-occurrence order,
    unique_map (input row of each voxel), inverse_map (voxel of each row), agree-or-ignore colabels.
    Returns, in order: coords[U,D], [features[U]], [colabels[U]], [unique_map], [inverse_map]."""
    is_np = isinstance(coordinates, np.ndarray)
    if not is_np and not isinstance(coordinates, torch.Tensor):
        raise ValueError("coordinates must be a numpy array or a torch tensor")
    assert coordinates.ndim == 2, "The coordinates must be a 2D matrix. The shape of the input is " + str(coordinates.shape)
    if features is not None:
        assert features.shape[0] == coordinates.shape[0]
    if labels is not None:
        assert labels.shape[0] == coordinates.shape[0]
        assert features is not None or return_maps_only or True
    D = coordinates.shape[1]
    assert D == 3, "only 3 spatial dimensions are on the LiDOG path"
    src_dev = None if is_np else coordinates.device
    dev = coordinates.device if (not is_np and coordinates.is_cuda) else _device(device)
    c = torch.from_numpy(np.ascontiguousarray(coordinates)) if is_np else coordinates
    c = c.to(dev)
    if quantization_size is not None:
        q4 = quantize_points(c if c.dtype == torch.float64 else c.to(torch.float32), quantization_size)
    else:
        q4 = torch.zeros((c.shape[0], 4), dtype=torch.int32, device=dev)
        q4[:, 1:] = torch.floor(c).to(torch.int32) if c.is_floating_point() else c.to(torch.int32)
    lab = None
    if labels is not None:
        lab = (torch.from_numpy(np.ascontiguousarray(labels)) if isinstance(labels, np.ndarray) else labels).to(dev)
    res = coords_unique(q4, 1, labels=lab, ignore_label=ignore_label)
    umap, inv = res["unique_map"], res["inverse_map"]

    def back(t, like=None):
        if is_np:
            a = t.cpu().numpy()
            return a if like is None else a.astype(like.dtype, copy=False)
        return t.to(src_dev)

    if return_maps_only:
        return (back(umap), back(inv)) if return_inverse else back(umap)
    out = [back(res["coords"][:, 1:].contiguous())]
    if features is not None:
        if is_np and isinstance(features, np.ndarray):
            out.append(features[umap.cpu().numpy()])
        else:
            f = features if isinstance(features, torch.Tensor) else torch.from_numpy(features)
            out.append(f[umap.to(f.device)])
    if labels is not None:
        out.append(back(res["colabels"], labels if isinstance(labels, np.ndarray) else None))
    if return_index:
        out.append(back(umap))
    if return_inverse:
        out.append(back(inv))
    return out[0] if len(out) == 1 else tuple(out)


def sparse_quantize_batch(points_list, labels_list, quantization_size, ignore_label=-100):
    """Voxelise a whole batch of device point clouds with one hash build.

    -> dict(coords int32 [U,4] (batch first), unique_map int64 [U] into the concatenated points,
            inverse_map int64 [N], colabels int32 [U], levels = every coordinate level of the batch
            (`CoordinateManager.from_quantized` adopts them: one hashed voxel index for voxelisation and network))."""
    dev = points_list[0].device
    sizes = [p.shape[0] for p in points_list]
    pts = torch.cat(points_list, 0)
    # batch index per point, built from fill kernels: torch.tensor(sizes, device=...) is a synchronous host-to-device
    # copy, i.e. a full stream sync at the start of every step (tools/host_profile.py)
    b = torch.cat([torch.full((n,), i, dtype=torch.int32, device=dev) for i, n in enumerate(sizes)], 0)
    q4 = quantize_points(pts, quantization_size, b)
    lab = None if labels_list is None else torch.cat(labels_list, 0)
    levels = build_levels(q4, labels=lab, ignore_label=ignore_label)  # stride 1 + the encoder's strides, one sync
    res = dict(levels[0])
    res["levels"] = levels
    return res


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """list of [N_b, 3] -> [sum N_b, 4] with the batch index in column 0."""
    rows = []
    for b, c in enumerate(coords):
        c = torch.from_numpy(c) if isinstance(c, np.ndarray) else c
        col = torch.full((c.shape[0], 1), b, dtype=c.dtype, device=c.device)
        rows.append(torch.cat([col, c], dim=1))
    out = torch.cat(rows, 0).to(dtype)
    return out if device is None else out.to(device)


class SparseCollation:
    """list of (coords, feats, labels) -> (batched coords in `dtype`, feats, labels), list order kept
    (utils/collation/collation.py:309-310)."""

    def __init__(self, limit_numpoints=-1, dtype=torch.int32, device=None):
        self.limit_numpoints, self.dtype, self.device = limit_numpoints, dtype, device

    def __call__(self, list_data):
        coords, feats, labels = list(zip(*list_data))
        as_t = lambda a: torch.from_numpy(a) if isinstance(a, np.ndarray) else a
        bc = batched_coordinates([as_t(c) for c in coords], dtype=self.dtype, device=self.device)
        return bc, torch.cat([as_t(f) for f in feats], 0), torch.cat([as_t(l) for l in labels], 0)


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """ME's kaiming init for (K, Cin, Cout) kernels: fan_in = Cin*K, fan_out = Cout*K
    (used with mode='fan_out' at utils/models/minkunet_bev.py:404)."""
    if tensor.dim() == 3:
        rf, cin, cout = tensor.shape
    elif tensor.dim() == 2:
        rf, (cin, cout) = 1, tensor.shape
    else:
        raise ValueError("kernel must be (K, Cin, Cout) or (Cin, Cout)")
    mode = mode.lower()
    if mode not in ("fan_in", "fan_out"):
        raise ValueError(f"Mode {mode} not supported")
    fan = (cin if mode == "fan_in" else cout) * rf
    std = torch.nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)
