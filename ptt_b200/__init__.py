"""ptt_b200 -- B200-native (sm_100a) point-feature hot path for PTT (shanjiayao/PTT).

Host side in Python/PyTorch (device memory, streams, torch.distributed only); the arithmetic is
hand-written CUDA behind the C ABI declared in include/ptt_b200.h.  Importing this package does
not load the native library; the first op call does, and raises if it is missing (no fallback).
"""
import os
import sys

__version__ = "0.1.0"

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install_dropin():
    """Make `import pointnet2_ops._ext` (pointnet2_utils.py:24 of the reference) resolve to the
    libptt_b200-backed module in ptt_b200/dropin, ahead of any other `pointnet2_ops` on sys.path."""
    if DROPIN_DIR in sys.path:
        sys.path.remove(DROPIN_DIR)
    sys.path.insert(0, DROPIN_DIR)
    for name in [m for m in sys.modules if m == "pointnet2_ops" or m.startswith("pointnet2_ops.")]:
        mod = sys.modules[name]
        if not getattr(mod, "__file__", "").startswith(DROPIN_DIR):
            del sys.modules[name]
    import pointnet2_ops._ext as ext

    return ext
