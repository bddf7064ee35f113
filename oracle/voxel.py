"""Oracle (test infrastructure): voxelisation, coordinate maps, kernel maps.

numpy restatement of the integer part of the hot path.  Conventions follow
SURVEY.md Appendix C; each function names the reference call site that fixes
its contract.  MinkowskiEngine 0.5.4 itself is absent => parity unpinned
(see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

_BIAS = 1 << 15  # 16 bits per axis, biased


def pack_keys(coords: np.ndarray) -> np.ndarray:
    """(b, x, y, z) or (x, y, z) int rows -> one int64 key per row (order-free)."""
    c = np.asarray(coords).astype(np.int64)
    if c.ndim != 2 or c.shape[1] not in (3, 4):
        raise ValueError("coords must be [N,3] or [N,4]")
    if c.shape[1] == 3:
        c = np.concatenate([np.zeros((c.shape[0], 1), np.int64), c], axis=1)
    sp = c[:, 1:]
    if c.shape[0] and (sp.min() < -_BIAS or sp.max() >= _BIAS or c[:, 0].min() < 0 or c[:, 0].max() >= (1 << 15)):
        raise ValueError("coordinate out of range for the packed 64-bit key")
    return (c[:, 0] << 48) | ((c[:, 1] + _BIAS) << 32) | ((c[:, 2] + _BIAS) << 16) | (c[:, 3] + _BIAS)


def unique_first_occurrence(coords: np.ndarray):
    """Unique rows in first-occurrence order.

    Returns (unique_map[U] ascending input rows, inverse_map[N] voxel id per row).
    Convention: SURVEY.md App. C.2 (`unique_map` indexes the *input* rows, as
    used by `points[voxel_idx]` at semantickitti_bev.py:240-242).
    """
    keys = pack_keys(coords)
    n = keys.shape[0]
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    _, first, inv_sorted = np.unique(keys, return_index=True, return_inverse=True)
    # np.unique orders by key; re-rank voxels by the row of their first occurrence
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    unique_map = first[order].astype(np.int64)
    inverse_map = rank[inv_sorted.reshape(-1)].astype(np.int64)
    return unique_map, inverse_map


def quantize_coords(points: np.ndarray, quantization_size) -> np.ndarray:
    """q = int32(floor(points / size)) in float32 -- or float64 for float64 points (App. C.1).

    numpy keeps `float32_array / python_float` in float32, which is what the
    call site semantickitti_bev.py:232-238 feeds ME; a length-3 size is applied
    per axis (minkunet_bev.py:279-284).
    """
    p = np.asarray(points)
    # float64 clouds (the reference's augmented training path: `coords @ R` with a float64 R,
    # utils/common/augmentation.py:10-20) stay float64 -- numpy divides an array by a python float in the ARRAY's
    # precision; every other dtype goes through float32 like the float32 clouds of the datasets
    dt = np.float64 if p.dtype == np.float64 else np.float32
    if p.dtype != dt:
        p = p.astype(dt)
    if np.isscalar(quantization_size):
        q = np.floor(p / dt(quantization_size))
    else:
        s = np.asarray(quantization_size, dtype=dt).reshape(1, -1)
        q = np.floor(p / s)
    return q.astype(np.int32)


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100,
                    return_index=False, return_inverse=False, return_maps_only=False,
                    quantization_size=None):
    """Restatement of ME.utils.sparse_quantize as LiDOG calls it.

    Call sites: semantickitti_bev.py:232-238 (5-tuple), mix3D.py:67-72 (4-tuple),
    minkunet_bev.py:279 (coords, index, inverse).  Return order (App. C.1-3):
    coords[U,D], [features[U]], [colabels[U]], [unique_map], [inverse_map].
    colabel = the voxel's label when all its points agree, else ignore_label.
    """
    coords = np.asarray(coordinates)
    assert coords.ndim == 2, "coordinates must be a 2D matrix"
    if features is not None:
        assert features.shape[0] == coords.shape[0]
    if labels is not None:
        assert labels.shape[0] == coords.shape[0]
    if quantization_size is not None:
        q = quantize_coords(coords, quantization_size)
    else:
        q = np.floor(coords).astype(np.int32)
    unique_map, inverse_map = unique_first_occurrence(q)
    if return_maps_only:
        return (unique_map, inverse_map) if return_inverse else unique_map
    out = [q[unique_map]]
    if features is not None:
        out.append(features[unique_map])
    if labels is not None:
        lab = np.asarray(labels)
        colabels = lab[unique_map].copy()
        disagree = lab != colabels[inverse_map]
        colabels[np.unique(inverse_map[disagree])] = ignore_label
        out.append(colabels)
    if return_index:
        out.append(unique_map)
    if return_inverse:
        out.append(inverse_map)
    return out[0] if len(out) == 1 else tuple(out)


def batched_coordinates(coords_list, dtype=np.int32) -> np.ndarray:
    """Prepend the batch index column, concatenate in list order (App. C.4;
    collation.py:309-325)."""
    rows = []
    for b, c in enumerate(coords_list):
        c = np.asarray(c)
        rows.append(np.concatenate([np.full((c.shape[0], 1), b, dtype=c.dtype), c], axis=1))
    if not rows:
        return np.zeros((0, 4), dtype)
    return np.concatenate(rows, axis=0).astype(dtype)


def stride_coords(coords: np.ndarray, new_stride: int):
    """Stride map (App. C.6): c' = floor(c / S) * S per spatial axis, batch kept;
    unique in first-occurrence order of the parent rows.

    Returns (coords_out[U,4] int32, parent_to_out[N] int64).
    """
    c = np.asarray(coords).astype(np.int64)
    s = int(new_stride)
    d = c.copy()
    d[:, 1:] = np.floor_divide(c[:, 1:], s) * s
    umap, inv = unique_first_occurrence(d)
    return d[umap].astype(np.int32), inv


def kernel_offsets(kernel_size: int, tensor_stride: int) -> np.ndarray:
    """Kernel offsets [K,3], x fastest (App. C.7).  Odd size is centred, even
    size starts at 0; scaled by the input tensor stride."""
    k = int(kernel_size)
    base = np.arange(k) - (k // 2 if k % 2 == 1 else 0)
    offs = np.zeros((k ** 3, 3), np.int64)
    i = 0
    for iz in range(k):
        for iy in range(k):
            for ix in range(k):
                offs[i] = (base[ix], base[iy], base[iz])
                i += 1
    return offs * int(tensor_stride)


def neighbor_table(in_coords: np.ndarray, out_coords: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """nbr[k, o] = row of (out_coords[o] + offsets[k]) in in_coords, else -1."""
    ci = np.asarray(in_coords).astype(np.int64)
    co = np.asarray(out_coords).astype(np.int64)
    in_keys = pack_keys(ci)
    order = np.argsort(in_keys, kind="stable")
    sorted_keys = in_keys[order]
    K = offsets.shape[0]
    nbr = np.full((K, co.shape[0]), -1, np.int64)
    if ci.shape[0] == 0 or co.shape[0] == 0:
        return nbr
    for k in range(K):
        q = co.copy()
        q[:, 1:] += offsets[k]
        ok = (q[:, 1:].min(axis=1) >= -_BIAS) & (q[:, 1:].max(axis=1) < _BIAS)
        qk = pack_keys(np.where(ok[:, None], q, co))
        pos = np.searchsorted(sorted_keys, qk)
        pos_c = np.minimum(pos, sorted_keys.shape[0] - 1)
        hit = ok & (sorted_keys[pos_c] == qk)
        nbr[k, hit] = order[pos_c[hit]]
    return nbr


def kernel_map(in_coords, out_coords, kernel_size: int, in_stride: int):
    """Forward kernel map: per offset k the (in_rows, out_rows) pair lists with
    out ascending.  Pair rule (App. C.7): out voxel o receives input at
    o + off_k through W[k]."""
    offs = kernel_offsets(kernel_size, in_stride)
    nbr = neighbor_table(in_coords, out_coords, offs)
    maps = []
    for k in range(offs.shape[0]):
        out_rows = np.nonzero(nbr[k] >= 0)[0].astype(np.int64)
        maps.append((nbr[k, out_rows].astype(np.int64), out_rows))
    return maps


def transposed_kernel_map(fine_coords, coarse_coords, kernel_size: int, fine_stride: int):
    """Transposed map (App. C.7): the forward (fine -> coarse) pair set with the
    roles swapped and the same k; in = coarse row, out = fine row.  Sorted by
    out within each k."""
    fwd = kernel_map(fine_coords, coarse_coords, kernel_size, fine_stride)
    maps = []
    for in_rows, out_rows in fwd:
        o = np.argsort(in_rows, kind="stable")
        maps.append((out_rows[o], in_rows[o]))
    return maps


def canonical_pairs(maps):
    """(k, out, in) triples sorted lexicographically (App. C.8)."""
    tr = []
    for k, (i, o) in enumerate(maps):
        tr.append(np.stack([np.full_like(o, k), o, i], axis=1))
    t = np.concatenate(tr, axis=0) if tr else np.zeros((0, 3), np.int64)
    if t.shape[0]:
        t = t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]
    return t
