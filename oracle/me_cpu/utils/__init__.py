"""Oracle stand-in for MinkowskiEngine.utils (sparse_quantize, SparseCollation,
kaiming_normal_); call sites: semantickitti_bev.py:232-238, collation.py:309,
minkunet_bev.py:404."""
import math

import numpy as np
import torch

from ... import voxel as _vox


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    is_t = isinstance(coordinates, torch.Tensor)
    c = coordinates.numpy() if is_t else coordinates
    f = features.numpy() if isinstance(features, torch.Tensor) else features
    l = labels.numpy() if isinstance(labels, torch.Tensor) else labels
    out = _vox.sparse_quantize(c, f, l, ignore_label, return_index, return_inverse, return_maps_only,
                               quantization_size)
    if is_t:
        out = tuple(torch.from_numpy(np.ascontiguousarray(o)) for o in out) if isinstance(out, tuple) \
            else torch.from_numpy(np.ascontiguousarray(out))
    return out


def batched_coordinates(coords, dtype=torch.int32, device=None):
    c = _vox.batched_coordinates([x.numpy() if isinstance(x, torch.Tensor) else x for x in coords])
    return torch.from_numpy(c).to(dtype)


class SparseCollation:
    """list of (coords, feats, labels) -> (coords[sumN,4] with batch column first, feats, labels)."""

    def __init__(self, limit_numpoints=-1, dtype=torch.int32, device=None):
        self.dtype, self.device = dtype, device

    def __call__(self, list_data):
        coords, feats, labels = list(zip(*list_data))
        as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))
        bc = batched_coordinates(coords, dtype=self.dtype)
        return bc, torch.cat([as_t(f) for f in feats], 0), torch.cat([as_t(l) for l in labels], 0)


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """fan_in = Cin * K, fan_out = Cout * K for a (K, Cin, Cout) kernel (App. C.9)."""
    if tensor.dim() == 3:
        rf, cin, cout = tensor.shape
    else:
        rf, (cin, cout) = 1, tensor.shape
    fan = (cin if mode == "fan_in" else cout) * rf
    std = torch.nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)


def sparse_quantize_batch(points_list, labels_list, quantization_size, ignore_label=-100):
    """Whole-batch voxelisation (same dict as the product's batched entry point): per-scan
    first-occurrence order, scans concatenated in list order, batch index in column 0."""
    coords, umaps, invs, colabs = [], [], [], []
    row0 = vox0 = 0
    for b, pts in enumerate(points_list):
        p = pts.numpy() if isinstance(pts, torch.Tensor) else np.asarray(pts)
        lab = None if labels_list is None else np.asarray(labels_list[b])
        if lab is None:
            q, um, inv = _vox.sparse_quantize(p, quantization_size=quantization_size, return_index=True,
                                              return_inverse=True)
            cl = None
        else:
            q, cl, um, inv = _vox.sparse_quantize(p, labels=lab, ignore_label=ignore_label,
                                                  quantization_size=quantization_size, return_index=True,
                                                  return_inverse=True)
        coords.append(np.concatenate([np.full((len(q), 1), b, np.int32), q], 1))
        umaps.append(um + row0)
        invs.append(inv + vox0)
        colabs.append(cl)
        row0 += len(p)
        vox0 += len(q)
    out = dict(coords=torch.from_numpy(np.concatenate(coords)), unique_map=torch.from_numpy(np.concatenate(umaps)),
               inverse_map=torch.from_numpy(np.concatenate(invs)), n=vox0,
               colabels=None if labels_list is None else torch.from_numpy(np.concatenate(colabs).astype(np.int32)))
    return out
