"""Host logic of the 16-bit gradient side channel (lidog_b200/me/_grad16.py) -- no GPU needed: a published copy is
handed out only for exactly the tensor it was made for, in the state it was published in; a gradient whose fp32
values were never written (LIDOG_BN_SKIP_DX32) can never be read as fp32, neither directly nor through the sum
autograd builds when the convolution output has a second consumer."""
import pytest
import torch

from lidog_b200.me import _grad16 as g


@pytest.fixture(autouse=True)
def _clean():
    g.clear()
    yield
    g.clear()


def _pub(valid):
    a = torch.zeros(4, 8)
    g.publish_grad16(a, a.half(), torch.ones(4), 1, fp32_valid=valid)
    return a


def test_hit_consumes_the_entry():
    a = _pub(False)
    g16, scale = g.take_grad16(a, 1)
    assert g16.dtype == torch.float16 and not g._TABLE and not g._NO_FP32
    assert g.take_grad16(a, 1) is None
    g.require_fp32(a, "t")  # nothing pending any more


def test_other_format_or_tensor_misses():
    a = _pub(True)
    assert g.take_grad16(torch.zeros(4, 8), 1) is None
    assert g.take_grad16(a, 2) is None and not g._TABLE  # fp32 values exist: entry dropped, the caller casts


def test_inplace_accumulation_invalidates_the_copy():
    a = _pub(True)
    a.add_(1.0)  # autograd accumulated a second gradient in place
    assert g.take_grad16(a, 1) is None
    g.require_fp32(a, "t")


def test_unwritten_fp32_is_never_readable():
    a = _pub(False)
    with pytest.raises(RuntimeError, match="LIDOG_BN_SKIP_DX32"):
        g.require_fp32(a, "t")
    s = torch.zeros(4, 8)  # out-of-place sum of `a` with another consumer's gradient
    assert g.take_grad16(s, 1) is None
    with pytest.raises(RuntimeError, match="LIDOG_BN_SKIP_DX32"):
        g.require_fp32(s, "t")
    a.add_(1.0)  # in-place sum
    assert g.take_grad16(a, 1) is None
    with pytest.raises(RuntimeError, match="LIDOG_BN_SKIP_DX32"):
        g.require_fp32(a, "t")
    g.require_fp32(torch.zeros(3, 8), "t")  # unrelated shapes are not affected
