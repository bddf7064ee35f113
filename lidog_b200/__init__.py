"""lidog_b200: B200-native (sm_100a) implementation of LiDOG's training hot path.

See DESIGN.md.  The CUDA library is loaded lazily by `lidog_b200.cabi`; the
Python layers under `lidog_b200.me` mirror the MinkowskiEngine API LiDOG calls.
"""
__version__ = "0.1.0"
