from lidog_b200.me.modules.resnet_block import BasicBlock, Bottleneck  # noqa: F401
