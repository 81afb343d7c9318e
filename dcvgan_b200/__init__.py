"""dcvgan_b200 - B200-native (sm_100a) implementation of the DCVGAN training step.

Host side: Python/PyTorch modules mirroring the reference surface (generator, discriminator, loss,
trainer).  Compute: hand-written CUDA kernels behind the C ABI of libdcvgan_b200.so
(include/dcvgan_b200.h).  There is no fallback path: without the built library and an sm_100
device every compute call raises.
"""
from ._lib import DcvError, lib, require_device  # noqa: F401

__version__ = "0.1.0"

_PRECISION = "bf16"


def set_precision(p):
    """'bf16' (tcgen05 fast path) or 'fp32' (exact CUDA-core path used for the 1e-3 parity gates)."""
    global _PRECISION
    assert p in ("bf16", "fp32")
    _PRECISION = p


def get_precision():
    return _PRECISION
