from . import resnet_block
