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
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        L = cabi.lib()
        dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = L.lg_peer_exchange_bytes()
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, pg.group_name)
        torch.cuda.synchronize()
        dist.barrier(group=pg)  # every buffer is zeroed before anyone raises a flag
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
    """The exchange of a process group, or None when disabled / unavailable (callers then use NCCL)."""
    if not CONFIG["enabled"]:
        return None
    key = id(pg)
    if key not in _INSTANCES:
        try:
            _INSTANCES[key] = PeerExchange(pg)
        except Exception as e:  # no symmetric memory on this system: NCCL carries the exchange
            warnings.warn(f"lidog_b200: peer-memory SyncBN exchange unavailable ({e!r}); using NCCL all_reduce")
            _INSTANCES[key] = None
    return _INSTANCES[key]
