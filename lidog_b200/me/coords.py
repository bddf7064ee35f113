"""CoordinateManager: the shared hashed voxel index of one batch.

Host-side mirror of MinkowskiEngine's coordinate manager as LiDOG exercises it
(ME.SparseTensor at utils/pipelines/trainer_lighting_2d.py:151; stride-2 maps and
kernel maps implied by the layers of utils/models/minkunet_bev.py:57-123).  It
owns, per tensor stride, the unique coordinate list and its GPU hash table, and
caches one gather plan per (stride_in, stride_out, kernel, kind).  All
arithmetic happens in liblidog_b200 (csrc/coords.cu, csrc/kmap.cu).
"""
from __future__ import annotations

import ctypes as C
import time

import torch

from .. import cabi
from . import _grad16


def _check_count(count_status: torch.Tensor, what: str) -> int:
    n, status = count_status.tolist()  # one host sync
    if status != 0:
        raise RuntimeError(f"{what}: coordinate outside the packed-key range "
                           f"(|x|,|y|,|z| < 32768, 0 <= batch < 32768); status {status}")
    return int(n)


def coords_unique(coords: torch.Tensor, stride: int = 1, labels: torch.Tensor | None = None, ignore_label: int = -100):
    """Wrapper of lg_coords_unique -> dict(coords, unique_map, inverse_map, colabels, table, capacity, n).
    One level, one host sync; the training path builds all its levels at once with `build_levels`."""
    assert coords.is_cuda and coords.dtype == torch.int32 and coords.dim() == 2 and coords.shape[1] == 4
    coords = coords.contiguous()
    L = cabi.lib()
    n = coords.shape[0]
    dev = coords.device
    cap = L.lg_hash_capacity(n)
    table = torch.empty(L.lg_hash_bytes(cap), dtype=torch.uint8, device=dev)
    out_coords = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    unique_map = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    inverse_map = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    colabels = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if labels is not None else None
    if labels is not None:
        labels = labels.to(device=dev, dtype=torch.int32).contiguous()
    count = torch.empty(2, dtype=torch.int64, device=dev)
    ws_bytes = L.lg_coords_unique_workspace(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cabi.check(L.lg_coords_unique(cabi.ptr(coords), n, stride, cabi.ptr(table), cap, cabi.ptr(out_coords),
                                  cabi.ptr(unique_map), cabi.ptr(inverse_map), cabi.ptr(labels), ignore_label,
                                  cabi.ptr(colabels), cabi.ptr(count), cabi.ptr(ws), ws_bytes, cabi.stream_of(coords)),
               "lg_coords_unique")
    nu = _check_count(count, "lg_coords_unique")
    return dict(coords=out_coords[:nu], unique_map=unique_map[:nu], inverse_map=inverse_map[:n],
                colabels=None if colabels is None else colabels[:nu], table=table, capacity=cap, n=nu)


# Tensor strides built together with the stride-1 level (MinkUNet34's encoder: minkunet_bev.py:62,69,76,83).
# Anything else is still derived on demand by `CoordinateManager.level` (one host round trip each).
PREBUILD_STRIDES = (2, 4, 8, 16)
_PINNED = {}
SYNC_WAIT = {"seconds": 0.0, "count": 0}  # host time blocked in the coordinate pyramid's synchronisation (bench.py)


def build_levels(coords: torch.Tensor, labels: torch.Tensor | None = None, ignore_label: int = -100,
                 strides=(1,) + PREBUILD_STRIDES):
    """All coordinate levels of a batch with ONE library call and ONE host synchronisation (lg_coords_pyramid): the
    row count of a level reaches the next one on the device.  Returns a list of per-level dicts like `coords_unique`
    (level 0 also carries colabels); `inverse_map` of level l > 0 maps the rows of level l-1 to level l."""
    assert coords.is_cuda and coords.dtype == torch.int32 and coords.dim() == 2 and coords.shape[1] == 4
    n = coords.shape[0]
    if n == 0:  # degenerate batch: the single-level call handles the empty case
        res = [coords_unique(coords, strides[0], labels, ignore_label)]
        for s in strides[1:]:
            res.append(coords_unique(res[-1]["coords"], s))
        return res
    coords = coords.contiguous()
    L = cabi.lib()
    dev = coords.device
    nl = len(strides)
    cap = L.lg_hash_capacity(n)
    tbytes = L.lg_hash_bytes(cap)
    if labels is not None:
        labels = labels.to(device=dev, dtype=torch.int32).contiguous()
    colabels = torch.empty(n, dtype=torch.int32, device=dev) if labels is not None else None
    arrays, outs = [], (cabi.LevelOut * nl)()
    for l in range(nl):
        table = torch.empty(tbytes, dtype=torch.uint8, device=dev)
        oc = torch.empty((n, 4), dtype=torch.int32, device=dev)
        maps = torch.empty((2, n), dtype=torch.int64, device=dev)
        arrays.append((table, oc, maps))
        outs[l] = cabi.LevelOut(table.data_ptr(), cap, oc.data_ptr(), maps[0].data_ptr(), maps[1].data_ptr())
    counts = torch.empty(2 * nl, dtype=torch.int64, device=dev)
    key = (dev.index, nl)
    host = _PINNED.get(key)
    if host is None:
        host = _PINNED[key] = torch.empty(2 * nl, dtype=torch.int64).pin_memory()
    stream = cabi.stream_of(coords)
    cabi.check(L.lg_coords_pyramid(coords.data_ptr(), n, cabi.ptr(labels), ignore_label, cabi.ptr(colabels), nl,
                                   (C.c_int32 * nl)(*strides), outs, counts.data_ptr(), host.data_ptr(), stream),
               "lg_coords_pyramid")
    cabi.count_launches("lg_coords_pyramid", 7 * nl + (1 if labels is not None else 0))
    t0 = time.perf_counter()
    torch.cuda.current_stream(dev).synchronize()  # the one host round trip of the step's coordinate work
    SYNC_WAIT["seconds"] += time.perf_counter() - t0
    SYNC_WAIT["count"] += 1
    cs = host.tolist()
    res, n_in = [], n
    for l in range(nl):
        nu, status = int(cs[2 * l]), cs[2 * l + 1]
        if status != 0:
            raise RuntimeError(f"lg_coords_pyramid: coordinate outside the packed-key range at stride {strides[l]} "
                               f"(|x|,|y|,|z| < 32768, 0 <= batch < 32768); status {status}")
        table, oc, maps = arrays[l]
        res.append(dict(coords=oc[:nu], unique_map=maps[0, :nu], inverse_map=maps[1, :n_in],
                        colabels=colabels[:nu] if (l == 0 and colabels is not None) else None, table=table,
                        capacity=cap, n=nu, stride=strides[l]))
        n_in = nu
    return res


class Level:
    """Coordinates of one tensor stride + their hash table.

    A level may be a contiguous row SLICE of a larger indexed set (`CoordinateManager.split`: several source batches
    voxelised and hashed together): `key_coords` are the rows as the shared table knows them (global batch index),
    `coords` what the tensor shows (`.C`, batch index local to the source), `row_offset` the global row of local row 0
    -- table hits and `parent_of_finer` entries are global and get the offset of their level subtracted."""

    def __init__(self, coords, table, capacity, parent_of_finer=None, key_coords=None, row_offset=0):
        self.coords, self.table, self.capacity = coords, table, capacity
        self.key_coords = coords if key_coords is None else key_coords
        self.row_offset = int(row_offset)
        self.n = coords.shape[0]
        self.parent_of_finer = parent_of_finer  # int64 [n_finer]: (global) row here of every row of the finer level


class GatherPlan:
    """Device arrays of one lgConvPlan (kept alive here) + the ctypes struct."""

    def __init__(self, nbr, k_stride, out_row, tile_mask, K, n_slots, n_out, n_in):
        self.nbr, self.out_row, self.tile_mask = nbr, out_row, tile_mask
        self.K, self.n_slots, self.n_out, self.n_in, self.k_stride = K, n_slots, n_out, n_in, k_stride
        self.c = cabi.make_plan(nbr, k_stride, out_row, tile_mask, K, n_slots, n_out, n_in)
        self.cref = C.byref(self.c)  # what the entry points take; built once, not per call
        self.n_tiles = n_slots // cabi.TILE
        self.key = None

    def count_pairs(self) -> int:
        """Number of (in, out) pairs of the map (host sync; used by the benchmark census only)."""
        return int((self.nbr >= 0).sum().item())


def _round_up(a, b):
    return (a + b - 1) // b * b


class CoordinateManager:
    def __init__(self, coordinates: torch.Tensor):
        _grad16.clear()  # a new batch: no gradient of the previous one is still wanted
        self.device = coordinates.device
        self._adopt(build_levels(coordinates), coordinates.shape[0])

    def _adopt(self, levels, n_input):
        first = levels[0]
        self.levels = {}
        for lv in levels:
            ts = lv.get("stride", 1)
            self.levels[ts] = Level(lv["coords"], lv["table"], lv["capacity"],
                                    parent_of_finer=None if lv is first else lv["inverse_map"])
        self.input_unique_map = first["unique_map"]
        self.input_inverse_map = first["inverse_map"]
        self.had_duplicates = first["n"] != n_input
        self.plans = {}

    @classmethod
    def from_quantized(cls, res):
        """Adopt the tables built by sparse_quantize_batch: voxelisation and the network share ONE hashed voxel
        index (no second hash build for ME.SparseTensor).  `res` = the level list of `build_levels`, or the dict of a
        single `coords_unique` call."""
        _grad16.clear()
        self = cls.__new__(cls)
        levels = res["levels"] if isinstance(res, dict) and "levels" in res else ([res] if isinstance(res, dict) else res)
        self.device = levels[0]["coords"].device
        self._adopt(levels, levels[0]["n"])
        return self

    def split(self, batch_sizes):
        """Per-source views of a manager built over the CONCATENATED batch of several sources (multi-source training,
        utils/pipelines/trainer_lighting_2d_multi.py:146-167: one forward pass per source through the same model).
        The hash tables, coordinate levels and one host round trip are shared; each view owns a contiguous row range
        of every level (rows are batch-ordered at every stride) and its own gather plans, built against the shared
        tables with the row offset of its slice.  -> list of CoordinateManager, one per entry of `batch_sizes`."""
        strides = sorted(self.levels)
        bounds = torch.tensor([sum(batch_sizes[:i]) for i in range(len(batch_sizes) + 1)], dtype=torch.int32,
                              device=self.device)
        cuts = torch.stack([torch.searchsorted(self.levels[ts].coords[:, 0].contiguous(), bounds) for ts in strides])
        cuts = cuts.tolist()  # one host round trip for all levels and sources
        views = []
        for s in range(len(batch_sizes)):
            v = CoordinateManager.__new__(CoordinateManager)
            v.device, v.levels, v.plans = self.device, {}, {}
            v.input_unique_map = v.input_inverse_map = None
            v.had_duplicates = False
            v.batch_size = int(batch_sizes[s])
            b0 = sum(batch_sizes[:s])
            for li, ts in enumerate(strides):
                lv = self.levels[ts]
                r0, r1 = cuts[li][s], cuts[li][s + 1]
                key = lv.key_coords[r0:r1]
                local = key.clone()
                if b0:
                    local[:, 0] -= b0
                parent = None
                if lv.parent_of_finer is not None:
                    f0, f1 = cuts[li - 1][s], cuts[li - 1][s + 1]
                    parent = lv.parent_of_finer[f0:f1]
                v.levels[ts] = Level(local, lv.table, lv.capacity, parent_of_finer=parent, key_coords=key,
                                     row_offset=lv.row_offset + r0)
            views.append(v)
        return views

    def _stream(self):
        return torch._C._cuda_getCurrentRawStream(self.device.index)

    # ------------------------------------------------------------------ coordinate levels
    def level(self, ts: int) -> Level:
        if ts not in self.levels:
            if ts < 2 or ts % 2:
                raise ValueError(f"tensor stride {ts} cannot be derived by stride-2 downsampling")
            fine = self.level(ts // 2)
            if fine.row_offset or fine.key_coords is not fine.coords:
                raise NotImplementedError("a split view only has the strides its parent manager was built with")
            res = coords_unique(fine.coords, ts)
            self.levels[ts] = Level(res["coords"], res["table"], res["capacity"], parent_of_finer=res["inverse_map"])
        return self.levels[ts]

    def get_coords(self, ts: int) -> torch.Tensor:
        return self.level(ts).coords

    # ------------------------------------------------------------------ gather plans
    def plan(self, kind: str, ts_in: int, ts_out: int, ksize: int) -> GatherPlan:
        key = (kind, ts_in, ts_out, ksize)
        if key not in self.plans:
            self.plans[key] = getattr(self, "_plan_" + kind)(ts_in, ts_out, ksize)
            self.plans[key].key = key
        return self.plans[key]

    def _neighbor_plan(self, lvl_in: Level, lvl_out: Level, ksize: int, scale: int) -> GatherPlan:
        L = cabi.lib()
        K = ksize ** 3
        n_out = lvl_out.n
        n_slots = _round_up(n_out, cabi.TILE)
        nbr = torch.empty((K, max(n_slots, 1)), dtype=torch.int32, device=self.device)
        mask = torch.empty((max(n_slots // cabi.TILE, 1), (K + 31) // 32), dtype=torch.int32, device=self.device)
        cabi.check(L.lg_kernel_map(cabi.ptr(lvl_in.table), lvl_in.capacity, cabi.ptr(lvl_out.key_coords), n_out, ksize,
                                   scale, lvl_in.row_offset, 1 if lvl_in is lvl_out else 0, cabi.ptr(nbr), n_slots,
                                   cabi.ptr(mask), self._stream()), "lg_kernel_map")
        return GatherPlan(nbr, n_slots, None, mask, K, n_slots, n_out, lvl_in.n)

    def _sorted_plan(self, lvl_in: Level, lvl_out: Level, ksize: int, scale: int) -> GatherPlan:
        """Mask-sorted processing order (lg_kernel_map_sorted): same pair set, rows with the same neighbour
        pattern share a tile.  This is what the convolution layers run on."""
        L = cabi.lib()
        K = ksize ** 3
        n_out = lvl_out.n
        n_slots = _round_up(n_out, cabi.TILE)
        nbr = torch.empty((K, max(n_slots, 1)), dtype=torch.int32, device=self.device)
        out_row = torch.empty(max(n_slots, 1), dtype=torch.int32, device=self.device)
        mask = torch.empty((max(n_slots // cabi.TILE, 1), 1), dtype=torch.int32, device=self.device)
        ws_bytes = L.lg_kernel_map_sorted_workspace(n_out, ksize)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        cabi.check(L.lg_kernel_map_sorted(cabi.ptr(lvl_in.table), lvl_in.capacity, cabi.ptr(lvl_out.key_coords), n_out,
                                          ksize, scale, lvl_in.row_offset, 1 if lvl_in is lvl_out else 0, cabi.ptr(nbr),
                                          cabi.ptr(out_row), n_slots, cabi.ptr(mask), cabi.ptr(ws), ws_bytes,
                                          self._stream()), "lg_kernel_map_sorted")
        return GatherPlan(nbr, n_slots, out_row, mask, K, n_slots, n_out, lvl_in.n)

    def _plan_same(self, ts_in, ts_out, ksize):
        assert ts_in == ts_out and ksize % 2 == 1
        lvl = self.level(ts_in)
        return self._neighbor_plan(lvl, lvl, ksize, ts_in)

    def _plan_down(self, ts_in, ts_out, ksize):
        assert ts_out == 2 * ts_in and ksize == 2
        return self._neighbor_plan(self.level(ts_in), self.level(ts_out), 2, ts_in)

    def _plan_same_sorted(self, ts_in, ts_out, ksize):
        assert ts_in == ts_out and ksize == 3
        lvl = self.level(ts_in)
        return self._sorted_plan(lvl, lvl, ksize, ts_in)

    def _plan_down_sorted(self, ts_in, ts_out, ksize):
        assert ts_out == 2 * ts_in and ksize == 2
        return self._sorted_plan(self.level(ts_in), self.level(ts_out), 2, ts_in)

    def _plan_up(self, ts_in, ts_out, ksize):
        """coarse (ts_in) -> fine (ts_out): fine rows grouped by child index, one k per tile."""
        assert ts_in == 2 * ts_out and ksize == 2
        L = cabi.lib()
        fine, coarse = self.level(ts_out), self.level(ts_in)
        n_fine = fine.n
        n_slots = _round_up(n_fine, cabi.TILE) + 8 * cabi.TILE
        gather = torch.empty(n_slots, dtype=torch.int32, device=self.device)
        out_row = torch.empty(n_slots, dtype=torch.int32, device=self.device)
        mask = torch.empty((n_slots // cabi.TILE, 1), dtype=torch.int32, device=self.device)
        used = torch.empty(1, dtype=torch.int64, device=self.device)
        ws_bytes = L.lg_scan_workspace(8 * n_fine) + 256
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        cabi.check(L.lg_kernel_map_up2(cabi.ptr(fine.coords), cabi.ptr(coarse.parent_of_finer), n_fine, ts_out,
                                       coarse.row_offset, cabi.ptr(gather), cabi.ptr(out_row), cabi.ptr(mask), n_slots,
                                       cabi.ptr(used), cabi.ptr(ws), ws_bytes, self._stream()), "lg_kernel_map_up2")
        return GatherPlan(gather, 0, out_row, mask, 8, n_slots, n_fine, coarse.n)

    def _plan_identity(self, ts_in, ts_out, ksize):
        assert ts_in == ts_out and ksize == 1
        n = self.level(ts_in).n
        n_slots = _round_up(n, cabi.TILE)
        nbr = torch.arange(n_slots, dtype=torch.int32, device=self.device)
        nbr[n:] = -1
        mask = torch.ones((max(n_slots // cabi.TILE, 1), 1), dtype=torch.int32, device=self.device)
        return GatherPlan(nbr, n_slots, None, mask, 1, n_slots, n, n)

    # ------------------------------------------------------------------ ME-format kernel maps (tests / interop)
    def kernel_map_pairs(self, plan: GatherPlan):
        """(in_rows, out_rows, k_offsets) sorted by (k, out) -- MinkowskiEngine's kernel-map format."""
        if plan.k_stride == 0 or plan.out_row is not None:
            raise ValueError("pair lists are defined for natural-order neighbour-table plans only")
        L = cabi.lib()
        total = plan.K * plan.n_slots
        in_rows = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        out_rows = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        k_off = torch.zeros(plan.K + 1, dtype=torch.int64, device=self.device)
        ws_bytes = L.lg_scan_workspace(total)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        cabi.check(L.lg_kernel_map_pairs(cabi.ptr(plan.nbr), plan.K, plan.n_slots, cabi.ptr(in_rows),
                                         cabi.ptr(out_rows), cabi.ptr(k_off), cabi.ptr(ws), ws_bytes, self._stream()),
                   "lg_kernel_map_pairs")
        k_off = k_off.cpu()
        p = int(k_off[-1])
        return in_rows[:p], out_rows[:p], k_off
