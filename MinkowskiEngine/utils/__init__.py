from lidog_b200.me.utils import *  # noqa: F401,F403
from lidog_b200.me.utils import sparse_quantize, SparseCollation, batched_coordinates, kaiming_normal_  # noqa: F401
