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
    """Precision of modules built from now on (and, for the fp32 storage modes, of every convolution launched from now on):
    'bf16' - bf16 activations, tcgen05 kind::f16, fp32 accumulation / statistics / master weights (fast path);
    'tf32' - fp32 activations; forward and data gradient of every convolution whose channel counts are multiples of 32 run
             on tcgen05 kind::tf32 (fp32 accumulation), the rest and the weight gradients on the CUDA-core kernels;
    'fp32' - fp32 activations, CUDA-core kernels only (exact mode)."""
    global _PRECISION
    assert p in ("bf16", "tf32", "fp32")
    _PRECISION = p
    from . import ops
    ops.TF32_TC = p == "tf32"


def get_precision():
    return _PRECISION
