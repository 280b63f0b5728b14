"""ptt_b200 -- B200-native (sm_100a) point-feature hot path for PTT (shanjiayao/PTT).

Host side in Python/PyTorch (device memory, streams, torch.distributed only); the arithmetic is
hand-written CUDA behind the C ABI declared in include/ptt_b200.h.  Importing this package does
not load the native library; the first op call does, and raises if it is missing (no fallback).
"""
__version__ = "0.1.0"
