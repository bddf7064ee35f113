"""GPU parity: the whole MinkUNet34BEV forward + backward against the CPU oracle with the same weights."""
import numpy as np
import pytest
import torch

from oracle import voxel as ov

from tests.helpers import record

pytestmark = pytest.mark.gpu


def _small_scan(seed=3):
    from lidog_b200.lidog import synth
    pts, lab = synth.make_scan(seed, "nuscenes")
    keep = (np.abs(pts[:, 0]) < 12) & (np.abs(pts[:, 1]) < 12)  # crop: keeps the CPU oracle fast
    return pts[keep], lab[keep]


def _build(ME, bev_fn, state=None, bound=12.0):
    from lidog_b200.lidog.model import MinkUNet34BEV
    torch.manual_seed(0)
    m = MinkUNet34BEV(1, 7, ME=ME, bev_fn=bev_fn, mapping_bound_2d=bound)
    if state is not None:
        m.load_state_dict(state)
    return m


@pytest.fixture(params=[False, True], ids=["head_fp32", "head_tf32"])
def head_tf32(request):
    """The dense 2D head stays in cuDNN (north_star) and PyTorch runs cuDNN convolutions in TF32 by default -- the
    precision the reference itself trains its head in.  `head_fp32` switches that off so the measured error is the
    sparse path's own; `head_tf32` is the configuration the bench (and the reference) run."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = request.param
    yield request.param
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("mode,tol", [("off", 2e-3), ("fp16", 3e-2)])
def test_model_forward_backward_matches_oracle(cuda, mode, tol, head_tf32):
    import MinkowskiEngine as ME
    from lidog_b200.me import conv as meconv
    from lidog_b200.lidog.bev import sparse2super
    from lidog_b200.lidog import losses
    from oracle import me_cpu
    from oracle.me_cpu.bevfn import sparse2super as o_s2s

    pts, lab = _small_scan()
    q, _, colab, umap, _ = ov.sparse_quantize(pts, np.ones((len(pts), 1), np.float32), lab, -1, True, True, False, 0.05)
    coords = ov.batched_coordinates([q])
    sem = torch.from_numpy(lab[umap]).long()

    ref_model = _build(me_cpu, o_s2s)
    state = ref_model.state_dict()
    xo = me_cpu.SparseTensor(coordinates=torch.from_numpy(coords), features=torch.ones(len(q), 1))
    out_o, bev_o = ref_model(xo, is_train=True)
    loss_o = losses.soft_dice_loss(out_o.F, sem, -1) + bev_o["block8"].square().mean()
    loss_o.backward()

    old = dict(meconv.CONFIG)
    meconv.CONFIG["tc"] = mode
    try:
        model = _build(ME, sparse2super, state).to(cuda)
        x = ME.SparseTensor(coordinates=torch.from_numpy(coords).to(cuda), features=torch.ones(len(q), 1, device=cuda))
        out, bev = model(x, is_train=True)
        loss = losses.soft_dice_loss(out.F, sem.to(cuda), -1) + bev["block8"].square().mean()
        loss.backward()
    finally:
        meconv.CONFIG.update(old)

    def rel(a, b):
        return float((a.detach().cpu().double() - b.detach().double()).norm() / b.detach().double().norm().clamp_min(1e-30))

    per_layer = {}
    ref_grads0 = dict(ref_model.named_parameters())
    for name, p in model.named_parameters():
        go = ref_grads0[name].grad
        if p.grad is not None and go is not None and float(go.norm()) > 1e-12:
            per_layer[name] = rel(p.grad, go)
    record("model_forward_backward", mode=mode, head_tf32=head_tf32, voxels=int(len(q)), logits=rel(out.F, out_o.F),
           bev=rel(bev["block8"], bev_o["block8"]), loss=abs(float(loss) - float(loss_o)),
           grad_worst=max(per_layer.values()), grad_median=float(np.median(list(per_layer.values()))),
           grad_worst_layer=max(per_layer, key=per_layer.get), per_layer=per_layer)
    assert rel(out.F, out_o.F) <= tol
    assert rel(bev["block8"], bev_o["block8"]) <= tol
    assert abs(float(loss) - float(loss_o)) <= tol * max(1.0, abs(float(loss_o)))
    ref_grads = dict(ref_model.named_parameters())
    worst = 0.0
    for name, p in model.named_parameters():
        g, go = p.grad, ref_grads[name].grad
        assert g is not None and go is not None, name
        if float(go.norm()) > 1e-12:
            worst = max(worst, rel(g, go))
    assert worst <= 20 * tol, worst  # gradients pass through 60 BN layers; looser than the per-layer bar
