"""BasicBlock / Bottleneck with ME's constructor keywords and sub-module names
(conv1/norm1/conv2/norm2/downsample), as used by utils/models/minkunet_bev.py:4,425-446;
the dataflow is the one the reference restates in utils/models/resnet_block.py:39-55,101-123."""
import torch.nn as nn

from ..conv import MinkowskiConvolution
from ..norm import MinkowskiBatchNorm, MinkowskiReLU


def _conv_bn(cin, cout, k, stride, dilation, momentum, dimension):
    return (MinkowskiConvolution(cin, cout, kernel_size=k, stride=stride, dilation=dilation, dimension=dimension),
            MinkowskiBatchNorm(cout, momentum=momentum))


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1, self.norm1 = _conv_bn(inplanes, planes, 3, stride, dilation, bn_momentum, dimension)
        self.conv2, self.norm2 = _conv_bn(planes, planes, 3, 1, dilation, bn_momentum, dimension)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.norm2(self.conv2(self.relu(self.norm1(self.conv1(x)))))
        out += x if self.downsample is None else self.downsample(x)
        return self.relu(out)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1, self.norm1 = _conv_bn(inplanes, planes, 1, 1, 1, bn_momentum, dimension)
        self.conv2, self.norm2 = _conv_bn(planes, planes, 3, stride, dilation, bn_momentum, dimension)
        self.conv3, self.norm3 = _conv_bn(planes, planes * self.expansion, 1, 1, 1, bn_momentum, dimension)
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.relu(self.norm2(self.conv2(out)))
        out = self.norm3(self.conv3(out))
        out += x if self.downsample is None else self.downsample(x)
        return self.relu(out)
