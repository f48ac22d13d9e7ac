"""tiledarray_b200 — B200-native contraction engine behind the TiledArray DistArray/expression API.

Only the contraction hot path is implemented; compute goes through libtadev.so
(hand-written sm_100a kernels + NCCL) and never through a CPU or PyTorch fallback.
"""
from ._lib import TadevError, OP_N, OP_T  # noqa: F401
from .device import Device, DeviceBuffer, device_count  # noqa: F401
