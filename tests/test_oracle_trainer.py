"""CPU: the trainer mirror on the ORACLE backend (test infrastructure): the multi-source step of
PLTTrainer2DMulti (utils/pipelines/trainer_lighting_2d_multi.py:135-199) equals the weighted sum of the two
single-source loss pairs, and one step moves the weights.  (The CUDA product runs the same Python on
`lidog_b200.me`; its parity with this backend is tests/test_gpu_trainer.py.)"""
import numpy as np
import torch

from lidog_b200.lidog import model as M, step, synth
from oracle import me_cpu
from oracle.me_cpu.bevfn import sparse2super as o_s2s


def _crop(seed, r):
    pts, lab = synth.make_scan(seed, "nuscenes")
    keep = (np.abs(pts[:, 0]) < r) & (np.abs(pts[:, 1]) < r)
    return torch.from_numpy(pts[keep]), torch.from_numpy(lab[keep])


def test_multi_source_step_is_the_weighted_sum_of_both_sources():
    torch.manual_seed(0)
    net = M.MinkUNet34BEV(1, 7, ME=me_cpu, bev_fn=o_s2s, mapping_bound_2d=30.0)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    tr = step.LidogTrainer(net, num_classes=7, shape="nuscenes", ME=me_cpu, source_weights=(0.5, 0.5))
    a, b = _crop(31, 5.0), _crop(32, 5.0)
    src0, src1 = ([a[0]], [a[1]]), ([b[0]], [b[1]])
    # expected value with the weights frozen (BN in training mode uses batch statistics: no dependence on order)
    want = 0.0
    for w, (p, l) in zip((0.5, 0.5), (src0, src1)):
        c, f, sem, bev, cm = tr.voxelize(p, l)
        _, l3, l2 = tr.forward_loss(c, f, sem, bev, len(p), cm)
        want += w * (float(l3) + float(l2))
    net.load_state_dict(before)  # the probing passes touched the BN running statistics
    got = float(tr.training_step_multi([src0, src1]))
    assert abs(got - want) < 1e-5 * max(1.0, abs(want)), (got, want)
    moved = max(float((v - before[k]).abs().max()) for k, v in net.state_dict().items() if v.dtype.is_floating_point)
    assert moved > 1e-4
