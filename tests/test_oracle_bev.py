"""CPU: the BEV oracle is pinned bit-exactly against golden vectors produced by the REFERENCE's own
sparse2super (tests/golden/make_bev_golden.py ran utils/models/minkunet_bev.py:169-230 on CPU)."""
import os

import numpy as np
import pytest

from oracle import bev as ob

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "bev_reference.npz"))


@pytest.mark.parametrize("name", ["small", "wide", "edge"])
def test_oracle_matches_reference_golden(name):
    coords, feats, bound = GOLD[f"{name}/coords"], GOLD[f"{name}/feats"], float(GOLD[f"{name}/bound"])
    B = int(coords[:, 0].max()) + 1
    out, ctx = ob.bev_forward(coords, feats, B, bound, policy="last")
    assert np.array_equal(out, GOLD[f"{name}/out"])
    assert np.array_equal(ob.bev_backward(GOLD[f"{name}/grad_out"], ctx), GOLD[f"{name}/grad_feats"])


def test_oracle_matches_reference_full_size():
    coords, feats = GOLD["full50/coords"], GOLD["full50/feats"]
    out, _ = ob.bev_forward(coords, feats, 2, 50.0)
    nz = np.nonzero(out.reshape(-1))[0]
    assert np.array_equal(nz, GOLD["full50/nz_index"]) and np.array_equal(out.reshape(-1)[nz], GOLD["full50/nz_value"])


def test_pixel_formula_is_not_the_integer_one():
    """the reference's three float32 roundings differ from c + H/2 for hundreds of columns."""
    c = np.arange(-1000, 1000)
    inb, px, py = ob.pixel_indices(np.stack([c, c, np.zeros_like(c)], 1), 50.0)
    assert inb[1:].all() and not inb[0]           # x = -50.0 is outside the strict bound
    assert (px[inb] != c[inb] + 1000).sum() > 100
    assert ob.image_size(50.0) == 2000 and ob.image_size(30.0) == 1200


def test_policies_agree_without_duplicates_and_differ_with():
    rng = np.random.default_rng(0)
    xy = rng.permutation(60 * 60)[:500]
    coords = np.stack([np.zeros(500), xy // 60 - 30, xy % 60 - 30, np.zeros(500)], 1).astype(np.int32)
    # distinct voxels can still share a pixel (float32 rounding of the pixel formula): keep one per pixel
    inb, px, py = ob.pixel_indices(coords[:, 1:], 2.0)
    _, keep = np.unique(py * 1000 + px, return_index=True)
    coords = coords[np.sort(keep)]
    n = len(coords)
    feats = rng.standard_normal((n, 4)).astype(np.float32)
    a, _ = ob.bev_forward(coords, feats, 1, 2.0, policy="last")
    b, _ = ob.bev_forward(coords, feats, 1, 2.0, policy="max")
    assert np.array_equal(a, b)  # one voxel per pixel: the policies coincide
    dup = coords.copy()
    dup[:, 3] = 1  # a second z-voxel on every pixel
    both = np.concatenate([coords, dup])
    a, ca = ob.bev_forward(both, np.concatenate([feats, feats + 1.0]), 1, 2.0, policy="last")
    b, cb = ob.bev_forward(both, np.concatenate([feats + 1.0, feats]), 1, 2.0, policy="max")
    assert np.array_equal(a, b)  # 'last' winner (second copy) == channel-wise max (first copy)
    g = np.ones_like(a)
    ga, gb = ob.bev_backward(g, ca), ob.bev_backward(g, cb)
    assert np.array_equal(ga[:n], ga[n:])            # 'last': losers receive the pixel's gradient too
    assert (gb[n:] == 0).all() and gb[:n].sum() > 0  # 'max': only the arg-max row
