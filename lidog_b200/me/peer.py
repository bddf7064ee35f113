"""Peer-memory exchange for MinkowskiSyncBatchNorm (csrc/peer.cu).

Each rank owns one small exchange buffer allocated as torch symmetric memory, i.e. mapped into every peer of
the node over NVLink.  `PeerExchange.sum` launches ONE kernel that publishes the local vector, raises an
epoch flag and adds the peers' vectors read through those mappings, in rank order -- it replaces the NCCL
all-reduce SyncBN issues per layer (train_lidog.py:228; 124 per step).  If symmetric memory cannot be set up
(single process, no peer access) callers fall back to `torch.distributed.all_reduce`.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

import torch

from .. import cabi

CONFIG = {"enabled": int(os.environ.get("LIDOG_PEER_SYNCBN", "1"))}
_INSTANCES = {}


class PeerExchange:
    def __init__(self, pg):
        import torch.distributed._symmetric_memory as symm_mem
        L = cabi.lib()
        dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = L.lg_peer_exchange_bytes()
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, pg.group_name)
        torch.cuda.synchronize()  # this rank's buffer is zeroed; `get` agrees with the peers before any flag is raised
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        self.ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        self.epoch = 0

    def sum(self, local: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """out = sum over ranks of `local` (float64 vectors of equal length on every rank)."""
        assert local.dtype == torch.float64 and out.dtype == torch.float64 and local.is_contiguous()
        self.epoch += 1
        cabi.check(cabi.lib().lg_peer_sum(cabi.ptr(local), local.numel(), self.ptrs, self.world, self.rank, self.epoch,
                                          cabi.ptr(out), cabi.stream()), "lg_peer_sum")
        return out


def get(pg):
    """The exchange of a process group, or None when disabled / unavailable (callers then use NCCL).

    The choice is COLLECTIVE: every rank tries the set-up, then the ranks agree with an all-reduce(MIN) of their
    success flags (which is also the barrier after which every buffer is known to be zeroed).  A rank whose
    rendezvous failed can therefore never leave its peers spinning on a flag it will not raise."""
    key = id(pg)
    if key not in _INSTANCES:
        import torch.distributed as dist
        ex, why = None, "disabled (LIDOG_PEER_SYNCBN=0)"
        if CONFIG["enabled"]:
            try:
                ex = PeerExchange(pg)
            except Exception as e:  # no symmetric memory on this system
                why = repr(e)
        ok = torch.tensor([1 if ex is not None else 0], dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=pg)
        if int(ok.item()) == 0:
            if CONFIG["enabled"]:
                warnings.warn(f"lidog_b200: peer-memory SyncBN exchange unavailable on at least one rank ({why}); "
                              f"using NCCL all_reduce")
            ex = None
        _INSTANCES[key] = ex
    return _INSTANCES[key]


def active(pg=None) -> bool:
    """Whether the peer-memory exchange (not NCCL) carries SyncBN for `pg` -- what bench.py reports."""
    if pg is None:
        import torch.distributed as dist
        pg = dist.group.WORLD
    return _INSTANCES.get(id(pg)) is not None
