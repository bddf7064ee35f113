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
    from lidog_b200.me import peer
    x2 = torch.randn(n, C, device=dev) - 0.5 * rank

    def run(fused, layer_calls, two_branches):
        norm.CONFIG["fused"], norm.CONFIG["layer_calls"] = fused, layer_calls
        dt = torch.float32 if fused else torch.float64
        ms = []
        for s in (1, 2):
            m = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(ME.MinkowskiBatchNorm(C)).to(dev).to(dt)
            with torch.no_grad():
                g = torch.Generator(device="cpu").manual_seed(s)
                m.bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
                m.bn.bias.copy_(torch.randn(C, generator=g) * 0.1)
            ms.append(m)
        xs, rs, x2s = (t.detach().clone().to(dt).requires_grad_(True) for t in (x, res, x2))
        o = ms[0](ME.SparseTensor(xs, coordinate_manager=cm))
        if two_branches:  # BasicBlock with a downsample branch: BN_a(x) + BN_b(x2)
            o += ms[1](ME.SparseTensor(x2s, coordinate_manager=cm))
        else:
            o += ME.SparseTensor(rs, coordinate_manager=cm)
        y = ME.MinkowskiReLU()(o).F
        y.backward(gy.to(dt))
        second = [x2s.grad, ms[1].bn.weight.grad, ms[1].bn.running_var] if two_branches else [rs.grad]
        return [y.detach(), xs.grad, ms[0].bn.weight.grad, ms[0].bn.bias.grad, ms[0].bn.running_mean,
                ms[0].bn.running_var] + second

    worst = 0.0
    report = {}
    for two in (False, True):
        ref = run(0, 1, two)
        for lc in (1, 0):  # one library call per layer with the in-kernel exchange / fine-grained calls + lg_peer_sum
            got = run(1, lc, two)
            errs = [rel(a, b) for a, b in zip(got, ref)]
            report[f"two_branches={two},layer_calls={lc}"] = f"{max(errs):.2e}"
            worst = max(worst, max(errs))
    norm.CONFIG["fused"], norm.CONFIG["layer_calls"] = 1, 1
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("syncbn fused vs torch SyncBatchNorm(f64), world", dist.get_world_size(), "exchange:",
              "peer memory (in-kernel)" if peer.active() else "NCCL", report, "worst over ranks", float(t))
        assert float(t) <= 1e-5, float(t)
        print("SYNCBN OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
