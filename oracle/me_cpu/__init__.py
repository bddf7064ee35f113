"""Oracle (test infrastructure): a CPU stand-in for the `MinkowskiEngine`
package surface LiDOG uses, built ONLY on the oracle's
numpy/torch-CPU arithmetic.  It is written independently of the CUDA product's
`lidog_b200.me` so the two can be compared, and it doubles as the CPU baseline
(`bench.py --impl reference` / `cpu_baseline`).  ME 0.5.4 is absent => parity
unpinned (oracle/__init__.py).
"""
from .core import (SparseTensor, CoordinateManager, MinkowskiConvolution, MinkowskiConvolutionTranspose,
                   MinkowskiBatchNorm, MinkowskiSyncBatchNorm, MinkowskiReLU, MinkowskiDropout, cat)
from . import utils, modules

__version__ = "0.5.4-oracle"
