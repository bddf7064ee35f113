"""GPU parity of one whole LiDOG training step (PLTTrainer2D.training_step, trainer_lighting_2d.py:141-293)
against the same step on the CPU oracle: batched voxelisation, BEV label image, MinkUNet34BEV, both DICE
losses, backward, Adam.  Inputs: a batch of two small scans -- one nuScenes-shaped crop and one Mix3D-shaped
merge (two crops re-quantised through float32 metres, utils/datasets/mix3D.py:43-87).

Tolerances (about twice the measured values, profiles/r02_*_parity_measured.jsonl): integer products of the data path
bit exact; losses within 2e-5 in both operand modes (measured 4e-6); parameter gradients, layer by layer, against the
float32 CPU oracle -- which is itself 2e-3 (median) / 7e-3 (worst layer) away from a float64 evaluation of this 63-layer
network (tests/test_gpu_model.py) -- within 3e-2 for the exact-fp32 kernels with the cuDNN head in true fp32
(measured 8e-3), 1e-1 with the head in PyTorch's default TF32 (4e-2), 5e-1 for fp16 tensor-core operands (2.3e-1).
"""
import numpy as np
import pytest
import torch

from tests.helpers import record

pytestmark = pytest.mark.gpu


def _crop(pts, lab, r):
    keep = (np.abs(pts[:, 0]) < r) & (np.abs(pts[:, 1]) < r)
    return pts[keep], lab[keep]


def _batch():
    from lidog_b200.lidog import synth
    a, la = _crop(*synth.make_scan(11, "nuscenes"), 9.0)
    # Mix3D-shaped: two crops, each snapped to its voxel grid, turned back into float32 metres, concatenated
    parts, labs = [], []
    for seed in (12, 13):
        p, l = _crop(*synth.make_scan(seed, "nuscenes"), 7.0)
        q, l1 = synth._quantize_first(p, l, 0.05)
        parts.append((q.astype(np.float32) * np.float32(0.05)).astype(np.float32))
        labs.append(l1)
    return [a, np.concatenate(parts)], [la, np.concatenate(labs).astype(np.int32)]


@pytest.fixture(params=[False, True], ids=["head_fp32", "head_tf32"])
def head_tf32(request):
    """The dense 2D head stays in cuDNN (north_star) and PyTorch runs cuDNN convolutions in TF32 by default -- the
    precision the reference itself trains its head in.  `head_fp32` switches that off so the measured error is the
    sparse path's own; `head_tf32` is the configuration the bench (and the reference) run."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = request.param
    yield request.param
    torch.backends.cudnn.allow_tf32 = old


LOSS_TOL = 2e-5
GRAD_TOL = {("off", False): 3e-2, ("off", True): 1e-1, ("fp16", False): 5e-1, ("fp16", True): 5e-1}


@pytest.mark.parametrize("mode", ["off", "fp16"])
def test_training_step_matches_oracle(cuda, mode, head_tf32):
    tol, gtol = LOSS_TOL, GRAD_TOL[(mode, head_tf32)]
    import MinkowskiEngine as ME
    from lidog_b200.me import conv as meconv
    from lidog_b200.lidog import model as M, step
    from oracle import me_cpu
    from oracle.me_cpu.bevfn import sparse2super as o_s2s

    pts, lab = _batch()
    torch.manual_seed(0)
    ref_net = M.MinkUNet34BEV(1, 7, ME=me_cpu, bev_fn=o_s2s, mapping_bound_2d=30.0)
    state = {k: v.clone() for k, v in ref_net.state_dict().items()}
    ref = step.LidogTrainer(ref_net, num_classes=7, shape="nuscenes", ME=me_cpu)
    P, L = [torch.from_numpy(p) for p in pts], [torch.from_numpy(l) for l in lab]
    # voxelisation + label products of the oracle trainer, then its loss terms
    c_o, f_o, sem_o, bev_o, cm_o = ref.voxelize(P, L)
    tot_o, l3_o, l2_o = ref.forward_loss(c_o, f_o, sem_o, bev_o, len(P), cm_o)
    ref.optimizer.zero_grad(set_to_none=True)
    tot_o.backward()
    ref.optimizer.step()

    old = dict(meconv.CONFIG)
    meconv.CONFIG["tc"] = mode
    try:
        net = M.MinkUNet34BEV(1, 7, mapping_bound_2d=30.0)
        net.load_state_dict(state)
        net = net.to(cuda)
        tr = step.LidogTrainer(net, num_classes=7, shape="nuscenes")
        Pd, Ld = [p.to(cuda) for p in P], [l.to(cuda) for l in L]
        c, f, sem, bev, cm = tr.voxelize(Pd, Ld)
        # integer products of the data path: bit exact
        assert torch.equal(c.cpu(), c_o.to(torch.int32)), "batched voxel coordinates differ"
        assert torch.equal(sem.cpu(), sem_o), "per-voxel 3D labels differ"
        assert torch.equal(bev.cpu(), bev_o), "BEV label images differ"
        tot, l3, l2 = tr.forward_loss(c, f, sem, bev, len(Pd), cm)
        tr.optimizer.zero_grad(set_to_none=True)
        tot.backward()
        tr.optimizer.step()
        torch.cuda.synchronize()
    finally:
        meconv.CONFIG.update(old)

    for name, a, b in (("3d", l3, l3_o), ("bev", l2, l2_o), ("total", tot, tot_o)):
        assert abs(float(a) - float(b)) <= tol * max(1.0, abs(float(b))), (name, float(a), float(b))
    # gradients, layer by layer (they pass through 62 BN layers: looser than the per-layer bar, as in test_gpu_model)
    def rel(a, b):
        return float((a.detach().cpu().double() - b.detach().double()).norm() / b.detach().double().norm().clamp_min(1e-30))

    ref_params = dict(ref_net.named_parameters())
    top = max(float(p.grad.norm()) for p in ref_params.values())
    bad = []
    measured = {}
    for name, p in net.named_parameters():
        if p.grad is not None and float(ref_params[name].grad.norm()) > 1e-6 * top:
            measured[name] = rel(p.grad, ref_params[name].grad)
    record("training_step", mode=mode, head_tf32=head_tf32, loss_3d=abs(float(l3) - float(l3_o)), loss_bev=abs(float(l2) - float(l2_o)),
           loss_total=abs(float(tot) - float(tot_o)), grad_worst=max(measured.values()),
           grad_median=float(np.median(list(measured.values()))), grad_worst_layer=max(measured, key=measured.get))
    for name, p in net.named_parameters():
        g, go = p.grad, ref_params[name].grad
        assert g is not None and go is not None and torch.isfinite(g).all(), name
        if float(go.norm()) > 1e-6 * top:  # gradients that are zero up to rounding (e.g. a bias in front of a BN) carry no signal
            e = rel(g, go)
            if e > gtol:
                bad.append((name, e, float(go.norm()), float(g.norm())))
    assert not bad, (top, bad[:8])
    # Adam moved the weights (|update| ~ lr on the first step) and nothing blew up
    moved = 0.0
    for name, p in net.named_parameters():
        assert torch.isfinite(p).all(), name
        moved = max(moved, float((p.detach().cpu() - state[name]).abs().max()))
    assert 0.5e-3 < moved < 2e-3, moved


def test_multi_source_step_on_one_shared_coordinate_manager_matches_oracle(cuda):
    """PLTTrainer2DMulti.training_step (trainer_lighting_2d_multi.py:135-199): one pass per source domain through the
    same model, here on views of ONE coordinate manager built over both batches (`LidogTrainer.voxelize_multi`),
    against the oracle trainer, which voxelises each source separately."""
    import MinkowskiEngine as ME
    from lidog_b200.lidog import model as M, step
    from oracle import me_cpu
    from oracle.me_cpu.bevfn import sparse2super as o_s2s
    pts, lab = _batch()
    sources = [([pts[0]], [lab[0]]), ([pts[1], pts[0][::2].copy()], [lab[1], lab[0][::2].copy()])]
    torch.manual_seed(0)
    ref_net = M.MinkUNet34BEV(1, 7, ME=me_cpu, bev_fn=o_s2s, mapping_bound_2d=30.0)
    state = {k: v.clone() for k, v in ref_net.state_dict().items()}
    ref = step.LidogTrainer(ref_net, num_classes=7, shape="nuscenes", ME=me_cpu)
    to = lambda src, dev: [([torch.from_numpy(p).to(dev) for p in P], [torch.from_numpy(l).to(dev) for l in Lb]) for P, Lb in src]
    loss_o = ref.training_step_multi(to(sources, "cpu"))
    net = M.MinkUNet34BEV(1, 7, mapping_bound_2d=30.0)
    net.load_state_dict(state)
    net = net.to(cuda)
    tr = step.LidogTrainer(net, num_classes=7, shape="nuscenes")
    batches = tr.voxelize_multi(to(sources, cuda))
    assert batches[0][4] is not batches[1][4] and batches[0][4].levels[1].table is batches[1][4].levels[1].table
    o_batches = [ref.voxelize(P, Lb) for P, Lb in to(sources, "cpu")]
    for (c, f, sem, bev, cm, nb), (co, fo, semo, bevo, cmo) in zip(batches, o_batches):
        assert torch.equal(c.cpu(), co.to(torch.int32)) and torch.equal(sem.cpu(), semo) and torch.equal(bev.cpu(), bevo)
    loss = tr.training_step_multi(to(sources, cuda))
    torch.cuda.synchronize()
    record("training_step_multi", loss=float(loss), loss_oracle=float(loss_o), diff=abs(float(loss) - float(loss_o)))
    assert abs(float(loss) - float(loss_o)) <= 1e-4 * max(1.0, abs(float(loss_o)))
    moved = max(float((p.detach().cpu() - state[k]).abs().max()) for k, p in net.named_parameters())
    assert 0.5e-3 < moved < 2e-3, moved
