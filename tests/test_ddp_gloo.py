"""CPU, world_size 2 (gloo): the data-parallel wiring of the training step -- whole scans per rank,
no data-path collective, gradient all-reduce through DistributedDataParallel -- exercised with the
CPU oracle layers (the CUDA layers need a GPU; the DDP/SyncBN wiring is the same nn.Module tree)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _small_trainer():
    from lidog_b200.lidog import model as M, step
    from oracle import me_cpu
    from oracle.me_cpu.bevfn import sparse2super
    torch.manual_seed(0)
    net = M.MinkUNet34BEV(1, 7, ME=me_cpu, bev_fn=sparse2super, mapping_bound_2d=6.0, layers=(1,) * 8)
    return net, step, me_cpu


def _scan(seed):
    from lidog_b200.lidog import synth
    pts, lab = synth.make_scan(seed, "nuscenes")
    keep = (np.abs(pts[:, 0]) < 6) & (np.abs(pts[:, 1]) < 6)
    return torch.from_numpy(pts[keep]), torch.from_numpy(lab[keep])


def _grads(net, step, me_cpu, seeds):
    tr = step.LidogTrainer(net, shape="nuscenes", ME=me_cpu)
    tr.bound, tr.bev_img = 6.0, 20
    scans = [_scan(s) for s in seeds]
    coords, feats, sem, bev, cm = tr.voxelize([p for p, _ in scans], [l for _, l in scans])
    total, _, _ = tr.forward_loss(coords, feats, sem, bev, len(scans), cm)
    net.zero_grad()
    total.backward()
    return {n: p.grad.clone() for n, p in net.named_parameters()}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    net, step, me_cpu = _small_trainer()
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    g = _grads(ddp, step, me_cpu, [100 + rank])  # each rank owns whole scans; seeds differ per rank
    g = {k.replace("module.", ""): v for k, v in g.items()}
    torch.save(g, os.path.join(out, f"g{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_two_ranks_average_gradients(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    net, step, me_cpu = _small_trainer()
    torch.set_num_threads(2)
    s0 = _grads(net, step, me_cpu, [100])
    s1 = _grads(net, step, me_cpu, [101])
    for k in s0:
        assert torch.equal(g0[k], g1[k]), k  # both ranks hold the all-reduced gradient
        ref = 0.5 * (s0[k] + s1[k])
        assert torch.allclose(g0[k], ref, rtol=1e-4, atol=1e-7), k


def test_sync_batchnorm_conversion_keeps_state():
    import MinkowskiEngine as ME
    from lidog_b200.lidog.model import MinkUNet34BEV
    m = MinkUNet34BEV(1, 7, layers=(1,) * 8)
    m.bn0.bn.running_mean.fill_(0.25)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    s = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(m)  # train_lidog.py:228
    after = s.state_dict()
    assert list(before) == list(after) and all(torch.equal(before[k], after[k]) for k in before)
    n_sync = sum(isinstance(x, ME.MinkowskiSyncBatchNorm) for x in s.modules())
    n_bn = sum(isinstance(x, ME.MinkowskiBatchNorm) for x in s.modules())
    assert n_sync == n_bn and n_sync > 20
    assert all(isinstance(x.bn, torch.nn.SyncBatchNorm) for x in s.modules() if isinstance(x, ME.MinkowskiSyncBatchNorm))
    # the dense 2D head keeps plain BatchNorm2d (only MinkowskiBatchNorm is converted)
    assert any(isinstance(x, torch.nn.BatchNorm2d) for x in s.modules())


def _flat_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lidog_b200.lidog import ddp as lddp
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.BatchNorm1d(8), torch.nn.Linear(8, 6), torch.nn.BatchNorm1d(6))
    before = {k: v.clone() for k, v in net.state_dict().items()}
    ddp, fb = lddp.wrap(net, buffers="flat")
    assert fb.intact() and len(fb.flat) == 2  # float32 statistics + int64 num_batches_tracked
    for k, v in net.state_dict().items():  # re-binding the buffers as views changed no value and no key
        assert torch.equal(v, before[k]), k
    ddp.train()
    torch.manual_seed(10 + rank)  # different data per rank -> un-synchronised BN running statistics diverge
    ddp(torch.randn(16, 4) * (1 + rank)).sum().backward()
    local = {k: v.clone() for k, v in net.named_buffers()}
    fb.broadcast()  # what DDP(broadcast_buffers=True) does at the start of the next forward
    torch.save({"local": local, "synced": {k: v.clone() for k, v in net.named_buffers()}}, os.path.join(out, f"b{rank}.pt"))
    assert fb.intact()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_buffer_broadcast_makes_rank0_authoritative(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_flat_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    b0, b1 = torch.load(tmp_path / "b0.pt"), torch.load(tmp_path / "b1.pt")
    assert any(not torch.equal(b0["local"][k], b1["local"][k]) for k in b0["local"])  # they did diverge
    for k in b0["local"]:
        assert torch.equal(b0["synced"][k], b0["local"][k]), k  # rank 0 keeps its own
        assert torch.equal(b1["synced"][k], b0["local"][k]), k  # rank 1 now holds rank 0's
