"""2+ rank check of the fused SyncBN path (torchrun, NCCL): fused MinkowskiSyncBatchNorm (+ReLU, residual) vs
torch SyncBatchNorm evaluated in float64 on the same per-rank data (ranks hold different row counts).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/syncbn_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def rel(a, b):
    b = b.double()
    return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import MinkowskiEngine as ME
    from lidog_b200.me import norm
    from tests.helpers import random_voxels
    coords = random_voxels(np.random.default_rng(10 + rank), 3000 + 700 * rank, span=40)
    cm = ME.SparseTensor(coordinates=torch.from_numpy(coords).to(dev),
                         features=torch.ones(len(coords), 1, device=dev)).coordinate_manager
    n, C = len(coords), 96
    torch.manual_seed(100 + rank)
    x = torch.randn(n, C, device=dev) * (1 + rank) + 0.3
    res = torch.randn(n, C, device=dev)
    gy = torch.randn(n, C, device=dev) * 1e-3
    worst = 0.0
    out = {}
    for fused in (1, 0):
        norm.CONFIG["fused"] = fused
        dt = torch.float32 if fused else torch.float64
        m = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(ME.MinkowskiBatchNorm(C)).to(dev).to(dt)
        with torch.no_grad():
            g = torch.Generator(device="cpu").manual_seed(1)
            m.bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
            m.bn.bias.copy_(torch.randn(C, generator=g) * 0.1)
        xs, rs = x.detach().clone().to(dt).requires_grad_(True), res.detach().clone().to(dt).requires_grad_(True)
        o = m(ME.SparseTensor(xs, coordinate_manager=cm))
        o += ME.SparseTensor(rs, coordinate_manager=cm)
        y = ME.MinkowskiReLU()(o).F
        y.backward(gy.to(dt))
        out[fused] = [y.detach(), xs.grad, rs.grad, m.bn.weight.grad, m.bn.bias.grad, m.bn.running_mean, m.bn.running_var]
    names = ["y", "dx", "dres", "dgamma", "dbeta", "running_mean", "running_var"]
    errs = {k: rel(a, b) for k, a, b in zip(names, out[1], out[0])}
    worst = max(errs.values())
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("syncbn fused vs torch SyncBatchNorm(f64):", {k: f"{v:.2e}" for k, v in errs.items()}, "worst over ranks", float(t))
        assert float(t) <= 1e-5, float(t)
        print("SYNCBN OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
