"""ctypes binding of libdcvgan_b200.so (the C ABI declared in include/dcvgan_b200.h).

The library is the only compute path of this package: there is no CPU or PyTorch fallback.
`lib()` raises if the shared object is missing; `require_device()` raises if the current CUDA
device is not an sm_100 part.
"""
import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libdcvgan_b200.so"
if os.environ.get("DCV_EXPERIMENTS_LIB", "0") == "1":      # tools/exp_*.py only: the -DDCV_EXPERIMENTS build (csrc/build.sh exp)
    LIB_PATH = _HERE / "libdcvgan_b200_exp.so"

DCV_F32, DCV_BF16 = 0, 1
ACT_NONE, ACT_LEAKY, ACT_TANH = 0, 1, 2
IMPL_SIMT, IMPL_TC, IMPL_TC_TF32 = 0, 1, 2
DIR_GATHER, DIR_SCATTER = 0, 1
LOSS_BCE_ONES, LOSS_BCE_ZEROS, LOSS_HINGE_REAL, LOSS_HINGE_FAKE, LOSS_SOFTPLUS_NEG = range(5)


class Geom(C.Structure):
    """mirror of `struct dcv_geom`"""

    _fields_ = [(n, C.c_int32) for n in (
        "N", "Tl", "Hl", "Wl", "Cl", "Ts", "Hs", "Ws", "Cs",
        "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw", "wCl", "wCs")]

    def key(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class PreBn(C.Structure):
    """mirror of `struct dcv_prebn` (BatchNorm + (Leaky)ReLU applied on load by the image-side kernels)"""

    _fields_ = [("mean", C.c_void_p), ("invstd", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("c0", C.c_int32), ("count", C.c_int32), ("slope", C.c_float)]


ABI_VERSION = 5   # include/dcvgan_b200.h DCV_ABI_VERSION

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_G = C.POINTER(Geom)

# name -> (restype, argtypes); must list every symbol of include/dcvgan_b200.h
SIGNATURES = {
    "dcv_abi_version": (_i, []),
    "dcv_last_error": (C.c_char_p, []),
    "dcv_device_ok": (_i, []),
    "dcv_launch_count": (C.c_longlong, []),
    "dcv_set_tuning": (_i, [C.c_char_p, _i]),
    "dcv_ingest_u8": (_i, [_i, _vp, _i64, _i, _vp, _i64, _vp]),
    "dcv_ingest_onehot": (_i, [_i, _vp, _i, _i64, _i, _vp, _i64, _vp]),
    "dcv_export_u8": (_i, [_i, _vp, _i64, _i, _i, _i, _i64, _vp, _vp]),
    "dcv_img_conv_supported": (_i, [_G, _i]),
    "dcv_img_conv_bwd_workspace_bytes": (_i64, [_G]),
    "dcv_img_conv_fwd": (_i, [_G, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i, _f, _vp]),
    "dcv_img_conv_scatter": (_i, [_G, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i, _f, _vp, _vp]),
    "dcv_img_conv_bwd": (_i, [_G, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i, _f, _vp, _i, _vp, _i64, _vp, _i64, _vp]),
    "dcv_packed_weight_bytes": (_i64, [_G, _i, _i]),
    "dcv_pack_weight": (_i, [_G, _i, _i, _vp, _i64, _i64, _i64, _vp, _vp]),
    "dcv_pack_weight_batch": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcv_pack_weight_multi": (_i, [_G, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcv_pack_weight_sub": (_i, [_G, _i, _i, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp]),
    "dcv_conv_tc_supported": (_i, [_G, _i]),
    "dcv_conv_tf32_supported": (_i, [_G, _i]),
    "dcv_head_loss_supported": (_i, [_G, _i]),
    "dcv_head_loss": (_i, [_G, _i, _vp, _i64, _vp, _vp, _i64, _i, _vp, _i, _vp, _i64, _f, _vp, _vp]),
    "dcv_conv": (_i, [_G, _i, _i, _i, _vp, _i64, _vp, _vp, _i64, _i, _f, _vp]),
    "dcv_conv_accumulate": (_i, [_G, _i, _vp, _i64, _vp, _vp, _i64, _vp]),
    "dcv_conv_stats_slots": (_i, [_G, _i, _i64, _i64]),
    "dcv_conv_stats": (_i, [_G, _i, _vp, _i64, _vp, _vp, _i64, _i, _f, _vp, _i, _vp]),
    "dcv_wgrad_workspace_bytes": (_i64, [_G, _i]),
    "dcv_wgrad_tc_supported": (_i, [_G]),
    "dcv_wgrad": (_i, [_G, _i, _i, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i, _vp, _i64, _vp]),
    "dcv_wgrad_partial": (_i, [_G, _i, _i, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "dcv_wgrad_reduce_sub": (_i, [_G, _i, _vp, _vp, _i64, _i64, _i64, _i, _i, _i, _i, _i, _vp]),
    "dcv_wgrad_reduce_multi": (_i, [_G, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcv_bn_stats_blocks": (_i, [_i64, _i]),
    "dcv_bn_stats": (_i, [_i, _vp, _i64, _i64, _i, _vp, _vp]),
    "dcv_bn_finalize": (_i, [_vp, _i, _i, _i64, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcv_bn_tail_workspace_bytes": (_i64, [_i64, _i]),
    "dcv_bn_tail_counters": (_i, []),
    "dcv_bn_stats_finalize": (_i, [_i, _vp, _i64, _i64, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcv_bn_act_bwd_reduce_finalize": (_i, [_i, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _f,
                                            _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "dcv_bn_eval_stats": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp]),
    "dcv_bn_act": (_i, [_i, _vp, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _vp, _i64, _vp]),
    "dcv_bn_act_bwd_reduce": (_i, [_i, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _vp, _vp]),
    "dcv_bn_bwd_finalize": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "dcv_bn_act_bwd_apply": (_i, [_i, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _vp,
                                  _i64, _vp, _i64, _vp]),
    "dcv_act_bwd": (_i, [_i, _vp, _i64, _vp, _i64, _i64, _i, _i, _f, _vp, _i64, _vp]),
    "dcv_col2im_act": (_i, [_i, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _i64, _i, _i, _vp]),
    "dcv_fold_w": (_i, [_i, _vp, _i64, _i, _vp, _vp, _i64, _i, _vp, _f, _i64, _i, _i, _i, _i, _vp, _i64, _vp]),
    "dcv_unfold_w": (_i, [_i, _vp, _i64, _i64, _i, _i, _i, _i, _vp, _i64, _i, _vp, _i64, _i, _vp]),
    "dcv_add_noise": (_i, [_i, _vp, _i64, _vp, _f, _i64, _i, _vp, _i64, _vp]),
    "dcv_axpy": (_i, [_i, _vp, _i64, _i64, _i, _vp, _i64, _i, _vp]),
    "dcv_tdiff": (_i, [_i, _vp, _i64, _i, _i, _i64, _i, _vp, _i64, _vp]),
    "dcv_tdiff_bwd": (_i, [_i, _vp, _i64, _i, _i, _i64, _i, _vp, _i64, _i, _vp]),
    "dcv_softmax": (_i, [_i, _vp, _i64, _i64, _i, _vp, _i64, _vp]),
    "dcv_softmax_bwd": (_i, [_i, _vp, _i64, _vp, _i64, _i64, _i, _vp, _i64, _vp]),
    "dcv_segm_remap": (_i, [_i, _vp, _i64, _i64, _i, _vp, _i64, _vp]),
    "dcv_to_channels_last": (_i, [_i, _vp, _i64, _i64, _i64, _i64, _i64, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    "dcv_from_channels_last": (_i, [_i, _vp, _i64, _i, _i, _i, _i, _i, _vp, _i64, _i64, _i64, _i64, _i64, _i, _vp]),
    "dcv_copy_cl": (_i, [_i, _vp, _i64, _i, _vp, _i64, _i64, _i, _vp]),
    "dcv_frame_copy": (_i, [_i, _vp, _i64, _i, _i, _i64, _i, _i, _vp, _vp, _i64, _i, _i, _vp]),
    "dcv_gru_traj_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "dcv_gru_traj_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "dcv_loss_fwd_bwd": (_i, [_i, _vp, _i64, _i64, _i, _vp, _i, _vp, _i64, _f, _vp]),
    "dcv_adam_multi": (_i, [_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64),
                            _f, _f, _f, _f, _f, _i64, _f, _vp]),
    "dcv_adam_flat_dev": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _vp, _f, _vp]),
    "dcv_adam_flat": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i64, _f, _vp]),
}

_lib = None


class DcvError(RuntimeError):
    pass


def lib():
    """Load the shared library once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DcvError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or dcvgan_b200/csrc/build.sh). dcvgan_b200 has no fallback compute path.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.dcv_abi_version() != ABI_VERSION:
            raise DcvError("libdcvgan_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise DcvError(lib().dcv_last_error().decode())


_device_checked = False


def require_device():
    """The CUDA path needs an sm_100 GPU; refuse to run anywhere else (no fallback)."""
    global _device_checked
    if _device_checked:
        return
    import torch
    if not torch.cuda.is_available():
        raise DcvError("dcvgan_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")
    if not lib().dcv_device_ok():
        raise DcvError(lib().dcv_last_error().decode())
    _device_checked = True
