"""Host logic of the dense-head patch (lidog_b200/lidog/bev.py:patch_reference_model, me/norm.py:double_conv_forward):
every module shaped like the reference's DoubleConv (utils/models/conv2d.py:9-25) gets the fused forward, nothing else
is touched, and off the GPU the patched forward is the module's own Sequential (same numbers, same state-dict keys)."""
import torch
import torch.nn as nn

from lidog_b200.lidog.bev import patch_reference_model


class _DoubleConvLike(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(cin, cout, 3, padding=1, stride=2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
            nn.Conv2d(cout, cout, 3, padding=1, stride=2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class _Other(nn.Module):
    def __init__(self):
        super().__init__()
        self.double_conv = nn.Sequential(nn.Conv2d(4, 4, 1), nn.ReLU())  # same attribute name, different structure

    def forward(self, x):
        return self.double_conv(x)


class _Model(nn.Module):
    mapping_bound_2d = 50.0

    def __init__(self):
        super().__init__()
        self.head = _DoubleConvLike(8, 16)
        self.other = _Other()


def test_patch_touches_only_double_conv_shaped_modules():
    torch.manual_seed(0)
    m = _Model()
    keys = list(m.state_dict().keys())
    x = torch.randn(2, 8, 13, 11)
    ref = m.head(x)
    m2 = patch_reference_model(m)
    assert m2 is m and list(m.state_dict().keys()) == keys
    assert "forward" in m.head.__dict__ and "forward" not in m.other.__dict__
    m.head.double_conv[1].reset_running_stats()
    m.head.double_conv[4].reset_running_stats()
    out = m.head(x)  # CPU tensor: the patched forward is the Sequential itself
    assert torch.equal(out, ref)
    m.eval()
    assert torch.equal(m.head(x), m.head.double_conv(x))


def test_mirror_head_keeps_reference_state_dict_keys():
    from lidog_b200.lidog.model import Encoder2D
    enc = Encoder2D(96, 7)
    keys = set(enc.state_dict().keys())
    for k in ("down1.maxpool_conv.0.double_conv.0.weight", "down1.maxpool_conv.0.double_conv.1.running_mean",
              "down1.maxpool_conv.0.double_conv.4.weight", "out_conv.conv.bias"):
        assert k in keys, k
    x = torch.randn(1, 96, 17, 17)
    assert enc(x).shape == (1, 7, 5, 5)
