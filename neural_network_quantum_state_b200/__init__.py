"""nqs-b200: B200-native (sm_100a) variational Monte Carlo for neural-network quantum states.

The product is the CUDA library libnqs_b200.so behind the C ABI of include/nqs_b200.h; this package is the thin Python
host side: `Engine` (ctypes over the ABI, numpy in/out), `sampler` (the pynqs-compatible surface), `dist` (one process per
GPU: chain sharding + NCCL bootstrap over torch.distributed)."""
from .engine import Engine, NQSError, SRResult  # noqa: F401

__all__ = ["Engine", "NQSError", "SRResult"]
