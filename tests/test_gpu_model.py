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


def _bev_any_dtype(x, bound=50.0, voxel_size=0.05, pool=(5, 3, 1), policy="last"):
    """sparse2super in the dtype of the features (the bit-exact oracle in oracle/bev.py is float32 only): the
    reference's own formulation -- index_put_ of the rows into a dense (H, W, C) image, raw re-view, max_pool2d
    (minkunet_bev.py:209-224) -- on the float32-rounded pixel indices of oracle.bev.pixel_indices, differentiated by
    torch.  Single-threaded so that the overwrite on duplicate pixels is 'highest row wins'."""
    from oracle import bev as ob
    coords = x.C.numpy()
    H = ob.image_size(bound, voxel_size)
    inb, px, py = ob.pixel_indices(coords[:, 1:], bound, voxel_size)
    outs = []
    torch.set_num_threads(1)
    for b in range(int(coords[:, 0].max()) + 1):
        rows = torch.from_numpy(np.nonzero(inb & (coords[:, 0] == b))[0])
        img = torch.zeros((H, H, x.F.shape[1]), dtype=x.F.dtype)
        img[torch.from_numpy(py)[rows], torch.from_numpy(px)[rows]] = x.F[rows]  # in place, like minkunet_bev.py:217
        outs.append(torch.nn.functional.max_pool2d(img.view(1, -1, H, H), pool[0], pool[1], pool[2]))
    return torch.cat(outs, 0)


@pytest.fixture(scope="module")
def oracle_runs():
    """The same forward + backward on the CPU oracle in float32 (the reference's precision) AND in float64 (ground
    truth).  The float64 run is what errors are measured against: a float32 evaluation of this 63-layer network is
    itself ~1e-2 away from it in the parameter gradients (62 batch norms in a row condition the backward badly on a
    22 k-voxel crop), so "GPU vs CPU-float32" mostly measures two float32 roundings against each other."""
    from lidog_b200.lidog import losses
    from oracle import me_cpu
    from oracle.me_cpu.bevfn import sparse2super as o_s2s
    pts, lab = _small_scan()
    q, _, colab, umap, _ = ov.sparse_quantize(pts, np.ones((len(pts), 1), np.float32), lab, -1, True, True, False, 0.05)
    coords = ov.batched_coordinates([q])
    sem = torch.from_numpy(lab[umap]).long()
    ref32 = _build(me_cpu, o_s2s)
    state = {k: v.clone() for k, v in ref32.state_dict().items()}
    runs = {}
    for name, model, dt in (("f32", ref32, torch.float32),
                            ("f64", _build(me_cpu, _bev_any_dtype, state).double(), torch.float64)):
        xo = me_cpu.SparseTensor(coordinates=torch.from_numpy(coords), features=torch.ones(len(q), 1, dtype=dt))
        out_o, bev_o = model(xo, is_train=True)
        loss_o = losses.soft_dice_loss(out_o.F, sem, -1) + bev_o["block8"].square().mean()
        loss_o.backward()
        runs[name] = dict(logits=out_o.F.detach(), bev=bev_o["block8"].detach(), loss=float(loss_o),
                          grads={k: p.grad.detach() for k, p in model.named_parameters()})
    return dict(q=q, coords=coords, sem=sem, state=state, **runs)


def _rel(a, b):
    return float((a.detach().cpu().double() - b.detach().double()).norm() / b.detach().double().norm().clamp_min(1e-30))


def _grad_errors(grads, ref):
    out = {}
    for name, go in ref.items():
        if float(go.norm()) > 1e-12:
            out[name] = _rel(grads[name], go)
    return out


# End-to-end bars = about twice the measured values (profiles/r02_*_parity_measured.jsonl, head in true fp32):
#   mode off (exact-fp32 SIMT convolutions): logits 2.3e-6; parameter gradients 6.1e-3 (median over layers) where the
#       CPU float32 oracle itself is 1.7e-3 from float64;
#   mode fp16 (tensor-core operands, unit round-off 2^-11 -- the precision class of TF32, which PyTorch uses by default
#       for the reference's own cuDNN head): logits 2.0e-3, BEV logits 1.4e-3 after 63 layers (per layer: 3e-4).
BARS = {"off": dict(logits=2e-5, bev=2e-5, loss=2e-5), "fp16": dict(logits=4e-3, bev=3e-3, loss=2e-5)}
BARS_TF32_HEAD_BEV = {"off": 1.5e-3, "fp16": 3e-3}  # BEV logits once cuDNN runs its convolutions in TF32 (measured 6.6e-4 / 1.4e-3)


@pytest.mark.parametrize("mode", ["off", "fp16"])
def test_model_forward_backward_matches_oracle(cuda, mode, head_tf32, oracle_runs):
    import MinkowskiEngine as ME
    from lidog_b200.me import conv as meconv
    from lidog_b200.lidog.bev import sparse2super
    from lidog_b200.lidog import losses
    o = oracle_runs
    q, coords, sem = o["q"], o["coords"], o["sem"]
    old = dict(meconv.CONFIG)
    meconv.CONFIG["tc"] = mode
    try:
        model = _build(ME, sparse2super, o["state"]).to(cuda)
        x = ME.SparseTensor(coordinates=torch.from_numpy(coords).to(cuda), features=torch.ones(len(q), 1, device=cuda))
        out, bev = model(x, is_train=True)
        loss = losses.soft_dice_loss(out.F, sem.to(cuda), -1) + bev["block8"].square().mean()
        loss.backward()
    finally:
        meconv.CONFIG.update(old)
    grads = {k: p.grad for k, p in model.named_parameters()}
    assert all(g is not None and torch.isfinite(g).all() for g in grads.values())
    e_gpu = _grad_errors(grads, o["f64"]["grads"])          # GPU vs float64 ground truth
    e_cpu = _grad_errors(o["f32"]["grads"], o["f64"]["grads"])  # what a float32 CPU evaluation is worth
    m = dict(logits=_rel(out.F, o["f64"]["logits"]), bev=_rel(bev["block8"], o["f64"]["bev"]),
             loss=abs(float(loss) - o["f64"]["loss"]), logits_vs_f32=_rel(out.F, o["f32"]["logits"]),
             grad_worst=max(e_gpu.values()), grad_median=float(np.median(list(e_gpu.values()))),
             grad_worst_layer=max(e_gpu, key=e_gpu.get), cpu_f32_grad_worst=max(e_cpu.values()),
             cpu_f32_grad_median=float(np.median(list(e_cpu.values()))))
    record("model_forward_backward", mode=mode, head_tf32=head_tf32, voxels=int(len(q)), per_layer=e_gpu, **m)
    bars = dict(BARS[mode])
    if head_tf32:
        bars["bev"] = BARS_TF32_HEAD_BEV[mode]
    for k, bar in bars.items():
        assert m[k] <= bar * (max(1.0, abs(o["f64"]["loss"])) if k == "loss" else 1.0), (mode, k, m[k], bar)
    if mode == "off" and not head_tf32:
        # the exact-fp32 path is float32 noise like a CPU float32 evaluation of the same graph, a few times larger
        # (measured 6.1e-3 median / 1.1e-2 worst layer against 1.7e-3 / 5.0e-3 for the CPU oracle in float32: the
        # reductions of the BN backward and the convolution wgrad run in a different, GPU-shaped order); bars ~2x
        assert m["grad_median"] <= 1.5e-2 and m["grad_worst"] <= 2.5e-2, m
    else:
        # fp16 operands / a TF32 head: the gradients stay finite and within the measured envelope (x2)
        assert m["grad_median"] <= 0.35 and m["grad_worst"] <= 0.6, m
