"""GPU parity: fused BEV projection vs the reference's golden vectors and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import bev as ob

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "bev_reference.npz"))


def _run(cuda, coords, feats, B, bound, policy, grad_out=None):
    from lidog_b200.lidog.bev import bev_project
    f = torch.from_numpy(feats).to(cuda).requires_grad_(True)
    out = bev_project(torch.from_numpy(coords).to(cuda), f, B, bound, 0.05, (5, 3, 1), policy)
    g = None
    if grad_out is not None:
        out.backward(torch.from_numpy(grad_out).to(cuda))
        g = f.grad.cpu().numpy()
    return out.detach().cpu().numpy(), g


@pytest.mark.parametrize("name", ["small", "wide", "edge"])
def test_bev_matches_reference_golden(cuda, name):
    """policy 'last' == the reference's sparse2super run single-threaded on CPU: forward bit-exact,
    gradient within 1e-6 relative (float32 summation order of <= 4 terms differs)."""
    coords, feats, bound = GOLD[f"{name}/coords"], GOLD[f"{name}/feats"], float(GOLD[f"{name}/bound"])
    B = int(coords[:, 0].max()) + 1
    out, g = _run(cuda, coords, feats, B, bound, "last", GOLD[f"{name}/grad_out"])
    assert np.array_equal(out, GOLD[f"{name}/out"])
    ref = GOLD[f"{name}/grad_feats"]
    assert np.abs(g - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    # the owner-computes backward sums a cell's windows in torch's CPU order: record whether the bits agree too
    from tests.helpers import record
    record("bev_backward_vs_reference", case=name, bit_exact=bool(np.array_equal(g, ref)),
           max_abs_diff=float(np.abs(g - ref).max()))
    _, g2 = _run(cuda, coords, feats, B, bound, "last", GOLD[f"{name}/grad_out"])
    assert np.array_equal(g, g2)  # atomic-free: run-to-run deterministic


def test_bev_full_size_image_golden(cuda):
    coords, feats = GOLD["full50/coords"], GOLD["full50/feats"]
    out, _ = _run(cuda, coords, feats, 2, 50.0, "last")
    assert tuple(out.shape) == tuple(GOLD["full50/out_shape"])
    nz = np.nonzero(out.reshape(-1))[0]
    assert np.array_equal(nz, GOLD["full50/nz_index"])
    assert np.array_equal(out.reshape(-1)[nz], GOLD["full50/nz_value"])


@pytest.mark.parametrize("policy", ["last", "max"])
def test_bev_matches_oracle_both_policies(cuda, policy):
    rng = np.random.default_rng(11)
    import tests.golden.make_bev_golden as mk
    coords, feats = mk.make_case(rng, 3000, 3, 4.0, 6, dup_frac=0.4)
    ref, ctx = ob.bev_forward(coords, feats, 3, 4.0, policy=policy)
    gw = rng.standard_normal(ref.shape).astype(np.float32)
    out, g = _run(cuda, coords, feats, 3, 4.0, policy, gw)
    assert np.array_equal(out, ref)
    gref = ob.bev_backward(gw, ctx)
    assert np.abs(g - gref).max() <= 1e-6 * max(1.0, np.abs(gref).max())


def test_bev_96_channels_kitti_shape(cuda):
    """The production shape (96 channels, 2000 x 2000 -> 666 x 666), checked through the oracle on one sample."""
    rng = np.random.default_rng(12)
    import tests.golden.make_bev_golden as mk
    coords, feats = mk.make_case(rng, 4000, 1, 50.0, 96, dup_frac=0.3)
    ref, _ = ob.bev_forward(coords, feats, 1, 50.0, policy="last")
    out, _ = _run(cuda, coords, feats, 1, 50.0, "last")
    assert out.shape == (1, 96, 666, 666) and np.array_equal(out, ref)


@pytest.mark.parametrize("policy", ["last", "max"])
def test_bev_channels_last_layout_is_the_same_tensor(cuda, policy):
    """layout 1 (NHWC / channels_last): same logical values and gradients as the NCHW layout, bit for bit in
    the forward; the backward consumes a channels_last gradient without a copy."""
    from lidog_b200.lidog.bev import bev_project
    rng = np.random.default_rng(13)
    import tests.golden.make_bev_golden as mk
    coords, feats = mk.make_case(rng, 5000, 2, 6.0, 40, dup_frac=0.4)  # 40 channels: a partial channel group
    c = torch.from_numpy(coords).to(cuda)
    outs, grads = [], []
    gw = None
    for cl in (False, True):
        f = torch.from_numpy(feats).to(cuda).requires_grad_(True)
        out = bev_project(c, f, 2, 6.0, 0.05, (5, 3, 1), policy, channels_last=cl)
        assert out.is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
        if gw is None:
            gw = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype(np.float32)).to(cuda)
        out.backward(gw.contiguous(memory_format=torch.channels_last) if cl else gw)
        outs.append(out.detach().cpu().numpy())
        grads.append(f.grad.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    assert np.abs(grads[0] - grads[1]).max() <= 1e-6 * max(1.0, np.abs(grads[0]).max())
    ref, ctx = ob.bev_forward(coords, feats, 2, 6.0, policy=policy)
    assert np.array_equal(outs[1], ref)
