"""Host layer of the drop-in `MinkowskiEngine` package (CPU-only checks, no compute calls): constructor contracts,
parameter shapes and names, argument validation and the loud failure without a CUDA device.  Reference call
sites: utils/models/minkunet_bev.py:57-126 (module ctors), :401-408 (`.kernel`, `.bn`),
utils/datasets/semantickitti_bev.py:232-238 (`sparse_quantize`), utils/collation/collation.py:309-325."""
import math

import numpy as np
import pytest
import torch

import MinkowskiEngine as ME


def test_convolution_parameter_shapes_follow_me():
    c3 = ME.MinkowskiConvolution(32, 64, kernel_size=3, dimension=3)
    assert tuple(c3.kernel.shape) == (27, 32, 64) and c3.bias is None
    c5 = ME.MinkowskiConvolution(1, 32, kernel_size=5, dimension=3)           # conv0p1s1, minkunet_bev.py:57
    assert tuple(c5.kernel.shape) == (125, 1, 32)
    c1 = ME.MinkowskiConvolution(96, 7, kernel_size=1, bias=True, dimension=3)  # final, minkunet_bev.py:118-123
    assert tuple(c1.kernel.shape) == (96, 7) and tuple(c1.bias.shape) == (1, 7)
    down = ME.MinkowskiConvolution(32, 32, kernel_size=2, stride=2, dimension=3)
    up = ME.MinkowskiConvolutionTranspose(256, 128, kernel_size=2, stride=2, dimension=3)
    assert tuple(down.kernel.shape) == (8, 32, 32) and tuple(up.kernel.shape) == (8, 256, 128)
    assert "kernel_size=[3, 3, 3]" in repr(c3) and "MinkowskiConvolutionTranspose" in repr(up)
    assert set(dict(c1.named_parameters())) == {"kernel", "bias"}


def test_convolution_initialisation_range():
    torch.manual_seed(0)
    conv = ME.MinkowskiConvolution(64, 128, kernel_size=3, dimension=3)
    bound = 1.0 / math.sqrt(64 * 27)  # ME default: uniform(+-1/sqrt(Cin * K))
    assert float(conv.kernel.abs().max()) <= bound and float(conv.kernel.abs().max()) > 0.9 * bound
    tr = ME.MinkowskiConvolutionTranspose(64, 32, kernel_size=2, stride=2, dimension=3)
    assert float(tr.kernel.abs().max()) <= 1.0 / math.sqrt(32 * 8)  # transposed: fan of the OUTPUT channels


@pytest.mark.parametrize("kwargs,exc", [
    (dict(kernel_size=3, dimension=2), ValueError),
    (dict(kernel_size=3, dilation=2, dimension=3), NotImplementedError),
    (dict(kernel_size=7, dimension=3), NotImplementedError),
    (dict(kernel_size=3, stride=2, dimension=3), NotImplementedError),
])
def test_convolution_rejects_what_is_off_the_path(kwargs, exc):
    with pytest.raises(exc):
        ME.MinkowskiConvolution(8, 8, **kwargs)


def test_kaiming_normal_fan_out_uses_kernel_volume():
    torch.manual_seed(0)
    conv = ME.MinkowskiConvolution(16, 256, kernel_size=3, dimension=3)
    ME.utils.kaiming_normal_(conv.kernel, mode="fan_out", nonlinearity="relu")  # minkunet_bev.py:404
    want = math.sqrt(2.0 / (256 * 27))
    assert abs(float(conv.kernel.std()) - want) < 0.05 * want


def test_batchnorm_wraps_a_torch_batchnorm1d():
    bn = ME.MinkowskiBatchNorm(96)
    assert isinstance(bn.bn, torch.nn.BatchNorm1d) and bn.bn.num_features == 96  # `.bn` is read at minkunet_bev.py:407
    assert set(dict(bn.named_parameters())) == {"bn.weight", "bn.bias"}
    sync = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(torch.nn.Sequential(bn, ME.MinkowskiReLU()))
    assert isinstance(sync[0], ME.MinkowskiSyncBatchNorm)
    assert torch.equal(sync[0].bn.weight, bn.bn.weight) and torch.equal(sync[0].bn.running_var, bn.bn.running_var)


def test_sparse_tensor_argument_validation():
    feats = torch.ones(4, 1)
    with pytest.raises(ValueError):
        ME.SparseTensor(features=np.ones((4, 1)), coordinates=torch.zeros(4, 4, dtype=torch.int32))
    with pytest.raises(ValueError):
        ME.SparseTensor(features=feats)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ME.SparseTensor(features=feats, coordinates=torch.zeros(4, 4, dtype=torch.int32))


def test_sparse_quantize_argument_validation():
    pts = np.random.default_rng(0).random((10, 3)).astype(np.float32)
    with pytest.raises(ValueError):
        ME.utils.sparse_quantize([[0, 0, 0]])
    with pytest.raises(AssertionError):
        ME.utils.sparse_quantize(pts.reshape(-1))
    with pytest.raises(AssertionError):
        ME.utils.sparse_quantize(pts, features=np.ones((9, 1), np.float32))
    with pytest.raises(AssertionError):
        ME.utils.sparse_quantize(pts, labels=np.zeros(3, np.int32))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            ME.utils.sparse_quantize(pts, quantization_size=0.05)


def test_collation_and_batched_coordinates():
    a = np.array([[1, 2, 3], [4, 5, 6]], np.int32)
    b = np.array([[7, 8, 9]], np.int32)
    bc = ME.utils.batched_coordinates([a, b])
    assert bc.dtype == torch.int32 and bc.tolist() == [[0, 1, 2, 3], [0, 4, 5, 6], [1, 7, 8, 9]]
    coll = ME.utils.SparseCollation(dtype=torch.float32)  # collation.py:309: coordinates are cast to float there
    coords, feats, labels = coll([(a, np.ones((2, 1), np.float32), np.array([3, 4])),
                                  (b, np.ones((1, 1), np.float32), np.array([5]))])
    assert coords.dtype == torch.float32 and coords[:, 0].tolist() == [0.0, 0.0, 1.0]
    assert feats.shape == (3, 1) and labels.tolist() == [3, 4, 5]
