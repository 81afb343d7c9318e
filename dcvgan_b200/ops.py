"""Thin torch-tensor wrappers over the C ABI (one function per dcv_* entry point family).

Activations are torch tensors of shape (N, T, H, W, C), channels innermost, T == 1 for the
2-D networks.  A tensor may be a channel-slice view of a wider buffer (that is how the
torch.cat calls of the reference disappear); `ld` is then the pixel stride of the parent.
PyTorch only provides memory and streams here - every arithmetic op is a libdcvgan_b200 kernel.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_TANH, DCV_BF16, DCV_F32, DIR_GATHER, DIR_SCATTER, IMPL_SIMT, IMPL_TC, IMPL_TC_TF32, Geom,
                   check, lib)

__all__ = ["ConvSpec", "Act"]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def dcv_dtype(t):
    if t.dtype == torch.float32:
        return DCV_F32
    if t.dtype == torch.bfloat16:
        return DCV_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def torch_dtype(precision):
    return torch.float32 if precision in ("fp32", "tf32") else torch.bfloat16


class Act:
    """Channels-last activation (N, T, H, W, C) with an explicit pixel stride `ld` (elements).

    `base` is a flat torch tensor that owns the memory, `off` the element offset of channel 0 of
    pixel 0.  `ch(c0, c1)` gives a channel slice that shares the parent's memory and stride.
    """

    __slots__ = ("base", "off", "n", "t", "h", "w", "c", "ld", "cp")

    def __init__(self, base, off, n, t, h, w, c, ld, cp=None):
        self.base, self.off, self.n, self.t, self.h, self.w, self.c, self.ld = base, off, n, t, h, w, c, ld
        self.cp = c if cp is None else cp   # channels that may be read: c real ones + zero padding up to cp

    @staticmethod
    def empty(n, t, h, w, c, dtype, ld=None, zero=False):
        """Dense buffer.  A channel count that is not a multiple of 16 (image-like tensors: 1, 2, 3, 25, 50, 266
        channels) gets zero-filled padding channels up to the next multiple of 16; nothing ever writes them, so the
        tensor-core kernels can read the padded width (see dcv_geom.wCl / wCs)."""
        cp = c
        if ld is None:
            ld = c
            if c % 16 != 0 and dtype == torch.bfloat16:   # the fp32 (CUDA-core) path gains nothing from padding
                ld = cp = (c + 15) // 16 * 16
                zero = True
        numel = max(n * t * h * w * ld, 1)
        if zero and ZERO_POOL is not None:
            base = ZERO_POOL.get(numel, dtype, c)
        else:
            base = (torch.zeros if zero else torch.empty)(numel, dtype=dtype, device="cuda")
        return Act(base, 0, n, t, h, w, c, ld, cp)

    def padded_to(self, k):
        """view whose channel count is k: the c real channels plus k - c of the zero padding channels"""
        assert self.c <= k <= self.cp, (self.c, k, self.cp)
        return self if k == self.c else Act(self.base, self.off, self.n, self.t, self.h, self.w, k, self.ld, self.cp)

    def rows2d(self):
        """torch view (rows, c) over the real channels (strided by ld); plumbing only"""
        return self.base.as_strided((self.rows, self.c), (self.ld, 1), self.off)

    @staticmethod
    def from_dense(x):
        """wrap a contiguous torch tensor already shaped (N,T,H,W,C)"""
        assert x.is_contiguous() and x.dim() == 5
        n, t, h, w, c = x.shape
        return Act(x.view(-1), 0, n, t, h, w, c, c)

    def ch(self, c0, c1):
        assert 0 <= c0 < c1 <= self.c
        return Act(self.base, self.off + c0, self.n, self.t, self.h, self.w, c1 - c0, self.ld)   # slices see no padding

    @property
    def ptr(self):
        return self.base.data_ptr() + self.off * self.base.element_size()

    @property
    def dtype(self):
        return self.base.dtype

    @property
    def device(self):
        return self.base.device

    @property
    def shape(self):
        return (self.n, self.t, self.h, self.w, self.c)

    @property
    def rows(self):
        return self.n * self.t * self.h * self.w

    @property
    def spatial(self):
        return (self.t, self.h, self.w)

    def reshape_nt(self, n, t):
        """reinterpret (N*T, 1, H, W) frames as (N, T, H, W) clips or back (same memory)"""
        assert n * t == self.n * self.t
        return Act(self.base, self.off, n, t, self.h, self.w, self.c, self.ld, self.cp)

    def torch(self):
        """strided torch view (N,T,H,W,C) of the same memory (tests / debugging)"""
        ld = self.ld
        return self.base.as_strided(self.shape, (self.t * self.h * self.w * ld, self.h * self.w * ld, self.w * ld, ld, 1),
                                    self.off)

    def like(self, c=None, dtype=None):
        return Act.empty(self.n, self.t, self.h, self.w, self.c if c is None else c, self.dtype if dtype is None else dtype)


class ZeroPool:
    """Zero-padded scratch buffers that survive from one training iteration to the next.

    A padded activation (1, 2, 3, 25 ... real channels in a 16-multiple pitch) must have zeros in its padding channels
    and nothing on the path ever writes them: producers store the real channels only, the tensor-core epilogues store
    exact zeros (zero weight rows).  So the 67 MB memsets of the image-sized buffers are needed once, not once per
    iteration: the fused step asks for its padded buffers in a fixed order and gets the same, already padded, buffer
    for the same request every iteration (and the CUDA graph of the step then contains no memset nodes for them).
    """

    def __init__(self):
        self.bufs = {}      # (numel, dtype, real channels) -> [tensors]
        self.cursor = {}

    def begin_step(self):
        self.cursor = {}

    def get(self, numel, dtype, c):
        # keyed by the real channel count too: a buffer is only ever reused for tensors with the same real channels, so its
        # padding channels stay exactly zero even when the request order differs between iterations (update gating)
        key = (numel, dtype, c)
        i = self.cursor.get(key, 0)
        self.cursor[key] = i + 1
        lst = self.bufs.setdefault(key, [])
        if i == len(lst):
            lst.append(torch.zeros(numel, dtype=dtype, device="cuda"))
        return lst[i]


ZERO_POOL = None   # set by the fused trainer step


def cl_view(a):
    """(ptr, ld, rows, C)"""
    return a.ptr, a.ld, a.rows, a.c


class ConvSpec:
    """Static description of one convolution layer of the reference networks.

    kind 'conv'  : nn.Conv2d / nn.Conv3d, weight (cout, cin, *k)      - forward = GATHER  (L = input)
    kind 'convT' : nn.ConvTranspose2d,    weight (cin, cout, kh, kw)  - forward = SCATTER (L = output)
    """

    def __init__(self, kind, cin, cout, k, s, p):
        assert kind in ("conv", "convT")
        self.kind, self.cin, self.cout = kind, cin, cout
        self.k, self.s, self.p = tuple(k), tuple(s), tuple(p)  # (t, h, w)
        self.taps = self.k[0] * self.k[1] * self.k[2]

    def out_spatial(self, in_spatial):
        if self.kind == "conv":
            return tuple((i + 2 * p - k) // s + 1 for i, k, s, p in zip(in_spatial, self.k, self.s, self.p))
        return tuple((i - 1) * s - 2 * p + k for i, k, s, p in zip(in_spatial, self.k, self.s, self.p))

    def geom(self, n, in_spatial, cin_p=None, cout_p=None):
        """cin_p / cout_p: channel counts of the activations including zero padding (default: no padding)"""
        out = self.out_spatial(in_spatial)
        cin_p = self.cin if cin_p is None else cin_p
        cout_p = self.cout if cout_p is None else cout_p
        if self.kind == "conv":
            l, s_, cl, cs, wl, ws = in_spatial, out, cin_p, cout_p, self.cin, self.cout
        else:
            l, s_, cl, cs, wl, ws = out, in_spatial, cout_p, cin_p, self.cout, self.cin
        return Geom(n, l[0], l[1], l[2], cl, s_[0], s_[1], s_[2], cs, *self.k, *self.s, *self.p, wl, ws)

    # element strides of the PyTorch master weight seen as w[cl, cs, tap]
    def weight_strides(self):
        if self.kind == "conv":      # (cout=cs, cin=cl, taps)
            return self.taps, self.cin * self.taps, 1
        return self.taps, self.cout * self.taps, 1  # convT: (cin=cs, cout=cl, taps)

    @property
    def fwd_dir(self):
        return DIR_GATHER if self.kind == "conv" else DIR_SCATTER

    @property
    def bwd_dir(self):
        return DIR_SCATTER if self.kind == "conv" else DIR_GATHER


def _tc_ok(t):
    ptr, ld, _, _ = cl_view(t)
    return t.dtype == torch.bfloat16 and ptr % 16 == 0 and ld % 8 == 0


_FORCE_SIMT = bool(int(os.environ.get("DCV_FORCE_SIMT", "0")))   # debugging aid: bf16 storage without tcgen05


TF32_TC = False     # dcvgan_b200.set_precision("tf32"): fp32 convolutions with channel counts % 32 == 0 run on tcgen05 kind::tf32
STRICT_TC = False   # set by the fused trainer at production widths: a bf16 convolution that cannot run on tcgen05 is an error


def choose_conv_impl(g, direction, x):
    if not _FORCE_SIMT and x.dtype == torch.bfloat16 and _tc_ok(x) and lib().dcv_conv_tc_supported(C.byref(g), direction):
        return IMPL_TC
    if STRICT_TC and x.dtype == torch.bfloat16 and not _FORCE_SIMT:
        raise _lib.DcvError(f"bf16 convolution {g.key()} dir {direction} (ptr % 16 = {x.ptr % 16}, ld {x.ld}) is not eligible for the "
                            "tcgen05 kernel; refusing to fall back to the CUDA-core kernel silently")
    if (TF32_TC and not _FORCE_SIMT and x.dtype == torch.float32 and x.ptr % 16 == 0 and x.ld % 4 == 0
            and lib().dcv_conv_tf32_supported(C.byref(g), direction)):
        return IMPL_TC_TF32
    return IMPL_SIMT


def pack_weight(spec, g, direction, impl, weight):
    nbytes = lib().dcv_packed_weight_bytes(C.byref(g), direction, impl)
    if nbytes < 0:
        check(-1)
    out = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    s_l, s_s, s_tap = spec.weight_strides()
    w = weight.detach()
    assert w.is_contiguous() and w.dtype == torch.float32
    check(lib().dcv_pack_weight(C.byref(g), direction, impl, w.data_ptr(), s_l, s_s, s_tap, out.data_ptr(), _stream()))
    return out


def conv_weight_part(weight, cl_off, cs_off, fold_k=None, fold_kw=1):
    """Window description of one master conv weight (cs_cnt, cl_cnt, *k) inside a merged / folded packed matrix.
    Plain: input channel c sits at activation channel cl_off + c, the weight's taps are the geometry's taps.
    w-folded (fold_k = k of fold_kw taps along w moved into the channels, ops.fold_w): the geometry has kw = 1, its tap
    index t' maps to the weight's tap t'*fold_kw + fold_k - i.e. a base offset of fold_k elements and a tap stride of
    fold_kw."""
    w = weight.detach() if weight.requires_grad else weight
    assert w.is_contiguous() and w.dtype == torch.float32
    cs_cnt, cl_cnt = w.shape[0], w.shape[1]
    taps_full = w[0, 0].numel()
    off = 0 if fold_k is None else fold_k
    s_tap = 1 if fold_k is None else fold_kw
    return (w, off, cl_off, cl_cnt, cs_off, cs_cnt, taps_full, cl_cnt * taps_full, s_tap)


def pack_weight_merged(g, direction, impl, parts):
    """One packed matrix out of several master weights that each own a window of it (the two stem convolutions of a
    discriminator run as ONE convolution over [xg | xc] -> [hc | hg], discriminator.py:79-90,121-124,180-193,225-228).
    parts: conv_weight_part(...) tuples; everything outside the windows is zero."""
    nbytes = lib().dcv_packed_weight_bytes(C.byref(g), direction, impl)
    if nbytes < 0:
        check(-1)
    out = torch.empty(nbytes, dtype=torch.uint8, device=parts[0][0].device)
    n = len(parts)
    if impl == IMPL_TC and n <= 16:      # one launch assembles the whole matrix
        check(lib().dcv_pack_weight_multi(
            C.byref(g), direction, n, (C.c_void_p * n)(*[p[0].data_ptr() + 4 * p[1] for p in parts]),
            (C.c_int64 * n)(*[p[6] for p in parts]), (C.c_int64 * n)(*[p[7] for p in parts]), (C.c_int64 * n)(*[p[8] for p in parts]),
            (C.c_int * n)(*[p[2] for p in parts]), (C.c_int * n)(*[p[3] for p in parts]), (C.c_int * n)(*[p[4] for p in parts]),
            (C.c_int * n)(*[p[5] for p in parts]), out.data_ptr(), _stream()))
        return out
    for i, (w, off, cl_off, cl_cnt, cs_off, cs_cnt, s_l, s_s, s_tap) in enumerate(parts):
        check(lib().dcv_pack_weight_sub(C.byref(g), direction, impl, w.data_ptr() + 4 * off, s_l, s_s, s_tap, cl_off, cl_cnt,
                                        cs_off, cs_cnt, int(i == 0), out.data_ptr(), _stream()))
    return out


def wgrad_merged(g, xl, xs, parts, impl=None):
    """Weight gradient of a merged convolution: one pass over the activations leaves the split partial sums in the
    workspace, then every master weight reduces its own window.  parts: (conv_weight_part(dw, ...), accumulate)"""
    impl = choose_wgrad_impl(g, xl, xs) if impl is None else impl
    if TRACE is not None:
        TRACE.append(("wgrad", g.key(), 0, impl, xl.ld, xs.ld, xl.c, xs.c))
    nbytes = lib().dcv_wgrad_workspace_bytes(C.byref(g), impl)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=xl.device)
    lp, ldl, _, _ = cl_view(xl)
    sp, lds, _, _ = cl_view(xs)
    check(lib().dcv_wgrad_partial(C.byref(g), impl, dcv_dtype(xl), lp, ldl, sp, lds, ws.data_ptr(), nbytes, _stream()))
    n = len(parts)
    if n > 16:
        for (dw, off, cl_off, cl_cnt, cs_off, cs_cnt, s_l, s_s, s_tap), acc in parts:
            check(lib().dcv_wgrad_reduce_sub(C.byref(g), impl, ws.data_ptr(), dw.data_ptr() + 4 * off, s_l, s_s, s_tap, cl_off,
                                             cl_cnt, cs_off, cs_cnt, int(acc), _stream()))
        return
    P = [p for p, _ in parts]
    check(lib().dcv_wgrad_reduce_multi(
        C.byref(g), impl, ws.data_ptr(), n, (C.c_void_p * n)(*[p[0].data_ptr() + 4 * p[1] for p in P]),
        (C.c_int64 * n)(*[p[6] for p in P]), (C.c_int64 * n)(*[p[7] for p in P]), (C.c_int64 * n)(*[p[8] for p in P]),
        (C.c_int * n)(*[p[2] for p in P]), (C.c_int * n)(*[p[3] for p in P]), (C.c_int * n)(*[p[4] for p in P]),
        (C.c_int * n)(*[p[5] for p in P]), (C.c_int * n)(*[int(a) for _, a in parts]), _stream()))


def pack_weight_batch(jobs):
    """jobs: [(spec, g, direction, weight)] (tcgen05 impl, whole weights) -> [packed uint8 tensors], ONE kernel launch"""
    n = len(jobs)
    outs = []
    for spec, g, direction, weight in jobs:
        nbytes = lib().dcv_packed_weight_bytes(C.byref(g), direction, IMPL_TC)
        if nbytes < 0:
            check(-1)
        outs.append(torch.empty(nbytes, dtype=torch.uint8, device=weight.device))
    gp = (C.POINTER(Geom) * n)(*[C.pointer(j[1]) for j in jobs])
    dirs = (C.c_int * n)(*[j[2] for j in jobs])
    ws = (C.c_void_p * n)(*[j[3].detach().data_ptr() for j in jobs])
    st = [j[0].weight_strides() for j in jobs]
    sl = (C.c_int64 * n)(*[x[0] for x in st])
    ss = (C.c_int64 * n)(*[x[1] for x in st])
    stp = (C.c_int64 * n)(*[x[2] for x in st])
    op = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    check(lib().dcv_pack_weight_batch(n, gp, dirs, ws, sl, ss, stp, op, _stream()))
    return outs


TRACE = None   # set to a list to record every conv / wgrad call (tools/layer_bench.py)


def conv(g, direction, impl, x, wp, y, act=ACT_NONE, slope=0.0):
    if TRACE is not None:
        TRACE.append(("conv", g.key(), direction, impl, x.ld, y.ld, x.c, y.c))
    xp, ldx, _, _ = cl_view(x)
    yp, ldy, _, _ = cl_view(y)
    assert x.dtype == y.dtype
    check(lib().dcv_conv(C.byref(g), direction, impl, dcv_dtype(x), xp, ldx, wp.data_ptr(), yp, ldy, act, slope, _stream()))


HEAD_LOSS = os.environ.get("DCV_NO_HEAD_LOSS", "0") != "1"
_HEAD_COUNTER = {}


def head_loss_ok(g, direction, impl, x):
    return (HEAD_LOSS and impl == IMPL_TC and x.dtype == torch.bfloat16 and x.ptr % 16 == 0 and x.ld % 8 == 0
            and bool(lib().dcv_head_loss_supported(C.byref(g), direction)))


def head_loss(g, direction, x, wp, y, kind, loss_out, accumulate, want_grad, grad_scale=1.0):
    """discriminator head + its loss term in one launch; returns dL/dlogits (Act shaped like y) or None"""
    if TRACE is not None:
        TRACE.append(("conv", g.key(), direction, IMPL_TC, x.ld, y.ld, x.c, y.c))
    cnt = _HEAD_COUNTER.get(x.device)
    if cnt is None:
        cnt = _HEAD_COUNTER[x.device] = torch.zeros(4, dtype=torch.int32, device=x.device)
    # ONE real channel in a zero-filled 16-channel pitch: the head's backward reads the padded width
    dy = Act.empty(y.n, y.t, y.h, y.w, 1, y.dtype) if want_grad else None
    check(lib().dcv_head_loss(C.byref(g), direction, x.ptr, x.ld, wp.data_ptr(), y.ptr, y.ld, kind, loss_out.data_ptr(), int(accumulate),
                              None if dy is None else dy.ptr, 1 if dy is None else dy.ld, grad_scale, cnt.data_ptr(), _stream()))
    return dy


CONV_ACCUMULATE = os.environ.get("DCV_NO_CONV_ACCUMULATE", "0") != "1"


def conv_accumulate_ok(g, direction, impl, x, y):
    """the tcgen05 convolution can add its result into y (TMA reduce-add epilogue): bf16, aligned, >= 16 output channels"""
    return (CONV_ACCUMULATE and impl == IMPL_TC and y.dtype == torch.bfloat16 and y.ptr % 16 == 0 and y.ld % 8 == 0 and y.c % 16 == 0
            and x.dtype == torch.bfloat16)


def conv_accumulate(g, direction, x, wp, y):
    """y += correlate(x, wp)"""
    if TRACE is not None:
        TRACE.append(("conv", g.key(), direction, IMPL_TC, x.ld, y.ld, x.c, y.c))
    check(lib().dcv_conv_accumulate(C.byref(g), direction, x.ptr, x.ld, wp.data_ptr(), y.ptr, y.ld, _stream()))


def conv_stats_slots(g, direction, x, y):
    """> 0: the convolution can accumulate the BatchNorm batch statistics of its output in its epilogue (that many slots)"""
    if _FORCE_SIMT or y.ptr % 16:
        return 0
    n = lib().dcv_conv_stats_slots(C.byref(g), direction, x.ld, y.ld)
    return max(n, 0)


def conv_stats(g, direction, x, wp, y, slots):
    """y = correlate(x, wp) (no activation) + partial sums for dcv_bn_finalize; returns the fp32 (slots, 2, C) partials"""
    if TRACE is not None:
        TRACE.append(("conv", g.key(), direction, IMPL_TC, x.ld, y.ld, x.c, y.c))
    partials = torch.empty((slots, 2, y.c), dtype=torch.float32, device=y.device)
    check(lib().dcv_conv_stats(C.byref(g), direction, x.ptr, x.ld, wp.data_ptr(), y.ptr, y.ld, ACT_NONE, 0.0,
                               partials.data_ptr(), slots, _stream()))
    return partials


def bn_finalize(partials, rows, eps, momentum, running_mean, running_var, num_batches_tracked=None):
    """mean / invstd (+ running statistics) from (nblk, 2, C) partial sums"""
    nblk, _, c = partials.shape
    mean = torch.empty(c, dtype=torch.float32, device=partials.device)
    invstd = torch.empty_like(mean)
    check(lib().dcv_bn_finalize(partials.data_ptr(), nblk, c, rows, eps, momentum, _p(running_mean), _p(running_var),
                                _p(num_batches_tracked), mean.data_ptr(), invstd.data_ptr(), _stream()))
    return mean, invstd


def choose_wgrad_impl(g, xl, xs):
    if not _FORCE_SIMT and xl.dtype == torch.bfloat16 and _tc_ok(xl) and _tc_ok(xs) and lib().dcv_wgrad_tc_supported(C.byref(g)):
        return IMPL_TC
    if STRICT_TC and xl.dtype == torch.bfloat16 and not _FORCE_SIMT:
        raise _lib.DcvError(f"bf16 weight gradient {g.key()} is not eligible for the tcgen05 kernel; refusing to fall back silently")
    return IMPL_SIMT


def wgrad(spec, g, xl, xs, dw, accumulate=False, impl=None):
    impl = choose_wgrad_impl(g, xl, xs) if impl is None else impl
    if TRACE is not None:
        TRACE.append(("wgrad", g.key(), 0, impl, xl.ld, xs.ld, xl.c, xs.c))
    nbytes = lib().dcv_wgrad_workspace_bytes(C.byref(g), impl)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=xl.device)
    lp, ldl, _, _ = cl_view(xl)
    sp, lds, _, _ = cl_view(xs)
    s_l, s_s, s_tap = spec.weight_strides()
    assert dw.is_contiguous() and dw.dtype == torch.float32
    check(lib().dcv_wgrad(C.byref(g), impl, dcv_dtype(xl), lp, ldl, sp, lds, dw.data_ptr(), s_l, s_s, s_tap,
                          int(accumulate), ws.data_ptr(), nbytes, _stream()))


# ------------------------------------------------------------------------------------ direct image-side convolution
IMG_CONV = os.environ.get("DCV_NO_IMG_CONV", "0") != "1"


IMG_FWD, IMG_BWD, IMG_WGRAD, IMG_SCATTER = 0, 1, 2, 3


def img_conv_ok(spec, g, what, small, big):
    """3x3 / stride 1 / pad 1 image-side layers whose small (L) side has <= 3 real channels (Inconv: generator.py:171-176,
    Outconv: generator.py:272-277) in bf16: HBM-bound mma.sync kernels instead of tensor-core tiles over 16 zero-padded
    channels.  what: IMG_FWD (L -> S pass), IMG_BWD (full Inconv backward), IMG_WGRAD (weight gradient only)."""
    return (IMG_CONV and not _FORCE_SIMT and small.dtype == torch.bfloat16 and big.dtype == torch.bfloat16
            and big.ptr % 16 == 0 and big.ld % 8 == 0 and spec.weight_strides()[2] == 1
            and bool(lib().dcv_img_conv_supported(C.byref(g), what)))


def img_conv_fwd(spec, g, x, weight, y, act, slope):
    """y (the 64 / 128-channel S tensor) = act(correlate(x (the small L tensor), w))"""
    if TRACE is not None:
        TRACE.append(("img_conv_fwd", g.key(), 0, -1, x.ld, y.ld, x.c, y.c))
    s_l, s_s, s_tap = spec.weight_strides()
    w = weight.detach()
    assert w.is_contiguous() and w.dtype == torch.float32
    check(lib().dcv_img_conv_fwd(C.byref(g), x.ptr, x.ld, w.data_ptr(), s_l, s_s, s_tap, y.ptr, y.ld, act, slope, _stream()))


def make_prebn(pre):
    """pre: None or dict(mean, invstd, gamma, beta, c0, slope) -> (ctypes struct or None, tensors to keep alive)"""
    if pre is None:
        return None, ()
    keep = (pre["mean"], pre["invstd"], pre["gamma"], pre["beta"])
    st = _lib.PreBn(pre["mean"].data_ptr(), pre["invstd"].data_ptr(), _p(pre["gamma"]), _p(pre["beta"]), int(pre["c0"]), 64,
                    float(pre["slope"]))
    return C.byref(st), keep + (st,)


def img_conv_scatter(spec, g, xb, weight, y, act, slope, pre=None):
    """y (the small L tensor) = act(transposed correlation of xb (the 64 / 128-channel S tensor) with w): Outconv forward.
    pre: the channels [c0, c0 + 64) of xb hold a pre-BatchNorm tensor, normalised + activated on load (dcv_prebn)"""
    if TRACE is not None:
        TRACE.append(("img_conv_scatter", g.key(), 0, -1, xb.ld, y.ld, xb.c, y.c))
    s_l, s_s, s_tap = spec.weight_strides()
    w = weight.detach()
    assert w.is_contiguous() and w.dtype == torch.float32
    pp, _keep = make_prebn(pre)
    check(lib().dcv_img_conv_scatter(C.byref(g), xb.ptr, xb.ld, w.data_ptr(), s_l, s_s, s_tap, y.ptr, y.ld, act, slope, pp, _stream()))


def img_conv_bwd(spec, g, da, a, x, weight, act, slope, dw, accumulate, dx):
    """dw (fp32, master layout, or None) and dx (Act for the small L tensor, or None) in one pass over da (gradient w.r.t.
    the activated S tensor) and a (the activated S tensor; None with ACT_NONE); x is the small L tensor"""
    if TRACE is not None:
        TRACE.append(("img_conv_bwd" if dx is not None else "img_conv_wgrad", g.key(), 0, -1, da.ld, x.ld, da.c, x.c))
    s_l, s_s, s_tap = spec.weight_strides()
    w = weight.detach()
    nbytes = lib().dcv_img_conv_bwd_workspace_bytes(C.byref(g))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=w.device)
    ap, lda = (da.ptr, da.ld) if a is None else (a.ptr, a.ld)
    check(lib().dcv_img_conv_bwd(C.byref(g), da.ptr, da.ld, ap, lda, x.ptr, x.ld, w.data_ptr(), s_l, s_s, s_tap, act, slope,
                                 _p(dw), int(accumulate), None if dx is None else dx.ptr, 0 if dx is None else dx.ld,
                                 ws.data_ptr(), nbytes, _stream()))


# ------------------------------------------------------------------------------------ BatchNorm & friends
def _p(t):
    return None if t is None else t.data_ptr()


BN_FUSED_FINALIZE = os.environ.get("DCV_NO_BN_TAIL", "0") != "1"   # statistics / backward sums finalized inside the reduction launch
_BN_COUNTERS = {}


def _bn_counters(device):
    """zero-initialised, self-resetting ticket buffer of the in-kernel finalize (one per device; calls on a stream are ordered)"""
    t = _BN_COUNTERS.get(device)
    if t is None:
        t = _BN_COUNTERS[device] = torch.zeros(lib().dcv_bn_tail_counters(), dtype=torch.int32, device=device)
    return t


def bn_batch_stats(z, eps, momentum, running_mean, running_var, num_batches_tracked=None):
    zp, ldz, rows, c = cl_view(z)
    if BN_FUSED_FINALIZE:
        mean = torch.empty(c, dtype=torch.float32, device=z.device)
        invstd = torch.empty_like(mean)
        ws = torch.empty(lib().dcv_bn_tail_workspace_bytes(rows, c), dtype=torch.uint8, device=z.device)
        check(lib().dcv_bn_stats_finalize(dcv_dtype(z), zp, ldz, rows, c, eps, momentum, _p(running_mean), _p(running_var),
                                          _p(num_batches_tracked), mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(),
                                          _bn_counters(z.device).data_ptr(), _stream()))
        return mean, invstd
    nblk = lib().dcv_bn_stats_blocks(rows, c)
    partials = torch.empty((nblk, 2, c), dtype=torch.float32, device=z.device)
    mean = torch.empty(c, dtype=torch.float32, device=z.device)
    invstd = torch.empty_like(mean)
    if num_batches_tracked is not None:
        assert num_batches_tracked.dtype == torch.int64 and num_batches_tracked.is_cuda
    check(lib().dcv_bn_stats(dcv_dtype(z), zp, ldz, rows, c, partials.data_ptr(), _stream()))
    check(lib().dcv_bn_finalize(partials.data_ptr(), nblk, c, rows, eps, momentum, _p(running_mean), _p(running_var),
                                _p(num_batches_tracked), mean.data_ptr(), invstd.data_ptr(), _stream()))
    return mean, invstd


def bn_eval_stats(running_mean, running_var, eps):
    c = running_mean.numel()
    mean = torch.empty(c, dtype=torch.float32, device=running_mean.device)
    invstd = torch.empty_like(mean)
    check(lib().dcv_bn_eval_stats(running_mean.data_ptr(), running_var.data_ptr(), c, eps, mean.data_ptr(),
                                  invstd.data_ptr(), _stream()))
    return mean, invstd


def bn_act(z, mean, invstd, gamma, beta, drop, act, slope, out):
    zp, ldz, rows, c = cl_view(z)
    op, ldo, _, _ = cl_view(out)
    rows_per_n = rows // z.n
    check(lib().dcv_bn_act(dcv_dtype(z), zp, ldz, rows, c, _p(mean), _p(invstd), _p(gamma), _p(beta), _p(drop),
                           rows_per_n, act, slope, op, ldo, _stream()))


def bn_act_bwd(da, a, z, mean, invstd, gamma, beta, drop, act, slope, dz, dgamma, dbeta, accumulate=False, batch_stats=True):
    """batch_stats=False: the forward normalised with the running statistics (eval mode), so mean / invstd are constants
    and dz = gamma * invstd * du without the two batch-statistic terms; dgamma / dbeta keep their formulas."""
    dap, ldda, rows, c = cl_view(da)
    ap, lda, _, _ = cl_view(a)
    zp, ldz, _, _ = cl_view(z)
    dzp, lddz, _, _ = cl_view(dz)
    rows_per_n = rows // z.n
    nblk = lib().dcv_bn_stats_blocks(rows, c)
    sums = torch.empty((2, c), dtype=torch.float32, device=z.device)
    dt = dcv_dtype(z)
    if BN_FUSED_FINALIZE:
        ws = torch.empty(lib().dcv_bn_tail_workspace_bytes(rows, c), dtype=torch.uint8, device=z.device)
        check(lib().dcv_bn_act_bwd_reduce_finalize(dt, dap, ldda, ap, lda, zp, ldz, rows, c, mean.data_ptr(), invstd.data_ptr(),
                                                   _p(gamma), _p(beta), _p(drop), rows_per_n, act, slope, sums.data_ptr(),
                                                   _p(dgamma), _p(dbeta), int(accumulate), ws.data_ptr(),
                                                   _bn_counters(z.device).data_ptr(), _stream()))
    else:
        partials = torch.empty((nblk, 2, c), dtype=torch.float32, device=z.device)
        check(lib().dcv_bn_act_bwd_reduce(dt, dap, ldda, ap, lda, zp, ldz, rows, c, mean.data_ptr(), invstd.data_ptr(),
                                          _p(gamma), _p(beta), _p(drop), rows_per_n, act, slope, partials.data_ptr(), _stream()))
        check(lib().dcv_bn_bwd_finalize(partials.data_ptr(), nblk, c, sums.data_ptr(), _p(dgamma), _p(dbeta),
                                        int(accumulate), _stream()))
    if not batch_stats:
        sums.zero_()
    check(lib().dcv_bn_act_bwd_apply(dt, dap, ldda, ap, lda, zp, ldz, rows, c, mean.data_ptr(), invstd.data_ptr(),
                                     _p(gamma), _p(beta), _p(drop), rows_per_n, act, slope, sums.data_ptr(), rows, dzp, lddz,
                                     _stream()))


def act_bwd(da, a, act, slope, dz):
    dap, ldda, rows, c = cl_view(da)
    ap, lda, _, _ = cl_view(a)
    dzp, lddz, _, _ = cl_view(dz)
    check(lib().dcv_act_bwd(dcv_dtype(a), dap, ldda, ap, lda, rows, c, act, slope, dzp, lddz, _stream()))


def add_noise(x, noise, sigma, out):
    xp, ldx, rows, c = cl_view(x)
    op, ldo, _, _ = cl_view(out)
    assert noise.dtype == torch.float32 and noise.is_contiguous() and noise.numel() == rows * c
    check(lib().dcv_add_noise(dcv_dtype(x), xp, ldx, noise.data_ptr(), sigma, rows, c, op, ldo, _stream()))


def pack_weight_strided(g, direction, impl, weight, s_l, s_s, s_tap):
    """dcv_pack_weight with explicit element strides of the master weight seen as w[cl, cs, tap]"""
    nbytes = lib().dcv_packed_weight_bytes(C.byref(g), direction, impl)
    if nbytes < 0:
        check(-1)
    out = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    w = weight.detach()
    assert w.is_contiguous() and w.dtype == torch.float32
    check(lib().dcv_pack_weight(C.byref(g), direction, impl, w.data_ptr(), s_l, s_s, s_tap, out.data_ptr(), _stream()))
    return out


def col2im_act(p_act, y, cout, kh, kw, stride, pad, act, slope):
    """y (N,1,Oh,Ow,cout) = act(col2im(P)), P (N,1,Ih,Iw,cout*kh*kw) from the tap-unrolled 1x1 convolution"""
    assert p_act.t == 1 and y.t == 1 and y.n == p_act.n
    check(lib().dcv_col2im_act(dcv_dtype(y), p_act.ptr, p_act.ld, p_act.n, p_act.h, p_act.w, cout, kh, kw, stride, pad, act, slope,
                               y.ptr, y.ld, y.h, y.w, _stream()))


def fold_w(xg, xc, out, kw, sw, pw, noise_g=None, noise_c=None, sigma=0.0):
    """out (N,T,H,Ow, kw*(cg+cc)) <- [xg | xc] with the kw taps along w moved into the channel dimension (+ Noise)"""
    assert xg.shape[:4] == xc.shape[:4] and out.c == kw * (xg.c + xc.c)
    lines, w = xg.n * xg.t * xg.h, xg.w
    assert out.rows == lines * ((w + 2 * pw - kw) // sw + 1)
    check(lib().dcv_fold_w(dcv_dtype(xg), xg.ptr, xg.ld, xg.c, _p(noise_g), xc.ptr, xc.ld, xc.c, _p(noise_c), float(sigma), lines, w,
                           kw, sw, pw, out.ptr, out.ld, _stream()))


def unfold_w(d2, dxg, dxc, kw, sw, pw):
    """adjoint of fold_w: dxg (N,T,H,W,cg), dxc (N,T,H,W,cc) <- d2 (N,T,H,Ow, kw*(cg+cc))"""
    lines, w = dxg.n * dxg.t * dxg.h, dxg.w
    check(lib().dcv_unfold_w(dcv_dtype(d2), d2.ptr, d2.ld, lines, w, kw, sw, pw, dxg.ptr, dxg.ld, dxg.c, dxc.ptr, dxc.ld, dxc.c,
                             _stream()))


def axpy(x, out, accumulate):
    xp, ldx, rows, c = cl_view(x)
    op, ldo, _, _ = cl_view(out)
    check(lib().dcv_axpy(dcv_dtype(x), xp, ldx, rows, c, op, ldo, int(accumulate), _stream()))


def tdiff(x, out):
    xp, ldx, _, c = cl_view(x)
    op, ldo, _, _ = cl_view(out)
    n, t, h, w, _ = x.shape
    check(lib().dcv_tdiff(dcv_dtype(x), xp, ldx, n, t, h * w, c, op, ldo, _stream()))


def tdiff_bwd(dy, dx, accumulate):
    dyp, lddy, _, c = cl_view(dy)
    dxp, lddx, _, _ = cl_view(dx)
    n, t, h, w, _ = dx.shape
    check(lib().dcv_tdiff_bwd(dcv_dtype(dx), dyp, lddy, n, t, h * w, c, dxp, lddx, int(accumulate), _stream()))


def softmax(z, y):
    zp, ldz, rows, c = cl_view(z)
    yp, ldy, _, _ = cl_view(y)
    check(lib().dcv_softmax(dcv_dtype(z), zp, ldz, rows, c, yp, ldy, _stream()))


def softmax_bwd(dy, y, dz):
    dyp, lddy, rows, c = cl_view(dy)
    yp, ldy, _, _ = cl_view(y)
    dzp, lddz, _, _ = cl_view(dz)
    check(lib().dcv_softmax_bwd(dcv_dtype(y), dyp, lddy, yp, ldy, rows, c, dzp, lddz, _stream()))


def segm_remap(x, y):
    xp, ldx, rows, c = cl_view(x)
    yp, ldy, _, _ = cl_view(y)
    check(lib().dcv_segm_remap(dcv_dtype(x), xp, ldx, rows, c, yp, ldy, _stream()))


def to_channels_last(src, dst):
    """src: fp32 (N,C,H,W) or (N,C,T,H,W) with arbitrary strides -> dst (N,T,H,W,C) channels-last."""
    assert src.dtype == torch.float32 and src.is_cuda
    if src.dim() == 4:
        src = src.unsqueeze(2)
    n, c, t, h, w = src.shape
    assert tuple(dst.shape) == (n, t, h, w, c), (src.shape, dst.shape)
    dp, ld, _, _ = cl_view(dst)
    sn, sc, st, sh, sw = src.stride()
    check(lib().dcv_to_channels_last(dcv_dtype(dst), src.data_ptr(), sn, sc, st, sh, sw, n, c, t, h, w, dp, ld, _stream()))


def from_channels_last(src, dst, accumulate=False):
    """src channels-last (N,T,H,W,C) -> dst fp32 (N,C,T,H,W) or (N,C,H,W) with arbitrary strides."""
    assert dst.dtype == torch.float32 and dst.is_cuda
    if dst.dim() == 4:
        dst = dst.unsqueeze(2)
    n, c, t, h, w = dst.shape
    assert tuple(src.shape) == (n, t, h, w, c), (src.shape, dst.shape)
    sp, ld, _, _ = cl_view(src)
    sn, sc, st, sh, sw = dst.stride()
    check(lib().dcv_from_channels_last(dcv_dtype(src), sp, ld, n, c, t, h, w, dst.data_ptr(), sn, sc, st, sh, sw,
                                       int(accumulate), _stream()))


def ingest_u8(src, dst):
    """uint8 frames (N,T,H,W,C) as stored on disk -> channels-last Act, x / 127.5 - 1 (dataset.py:128-131,157-167)"""
    assert src.dtype == torch.uint8 and src.is_cuda and src.is_contiguous() and tuple(src.shape) == dst.shape, (src.shape, dst.shape)
    check(lib().dcv_ingest_u8(dcv_dtype(dst), src.data_ptr(), dst.rows, dst.c, dst.ptr, dst.ld, _stream()))


def ingest_onehot(idx, dst):
    """class indices (N,T,H,W) uint8 / int64 -> one-hot channels-last Act (dataset.py:177-181)"""
    assert idx.dtype in (torch.uint8, torch.int64) and idx.is_cuda and idx.is_contiguous() and tuple(idx.shape) == dst.shape[:4]
    check(lib().dcv_ingest_onehot(dcv_dtype(dst), idx.data_ptr(), idx.element_size(), dst.rows, dst.c, dst.ptr, dst.ld, _stream()))


def export_u8(src, dst):
    """channels-last Act (N,T,H,W,C) in [-1,1] -> uint8 (N,C,T,H,W) like util.videos_to_numpy (util.py:74-79)"""
    assert dst.dtype == torch.uint8 and dst.is_cuda and dst.is_contiguous() and tuple(dst.shape) == (src.n, src.c, src.t, src.h, src.w)
    check(lib().dcv_export_u8(dcv_dtype(src), src.ptr, src.ld, src.n, src.c, src.t, src.h * src.w, dst.data_ptr(), _stream()))


def _frame_index(t):
    """host int, or a device int32 tensor (CUDA-graph replay keeps the index on the device)"""
    if isinstance(t, torch.Tensor):
        assert t.dtype == torch.int32 and t.is_cuda
        return 0, t.data_ptr()
    return int(t), None


def frame_extract(src, t, dst):
    """dst (N,1,H,W,C) = frame t of src (N,T,H,W,C)   (the x[:, :, t] slice at trainer.py:299,307,347)"""
    sp, lds, _, c = cl_view(src)
    dp, ldd, _, _ = cl_view(dst)
    ti, tp = _frame_index(t)
    check(lib().dcv_frame_copy(dcv_dtype(src), sp, lds, src.n, src.t, src.h * src.w, c, ti, tp, dp, ldd, 0, 0, _stream()))


def frame_scatter(dsrc, t, dframe, accumulate=True):
    """adjoint of frame_extract: dsrc[:, t] (+)= dframe"""
    sp, lds, _, c = cl_view(dsrc)
    dp, ldd, _, _ = cl_view(dframe)
    ti, tp = _frame_index(t)
    check(lib().dcv_frame_copy(dcv_dtype(dsrc), sp, lds, dsrc.n, dsrc.t, dsrc.h * dsrc.w, c, ti, tp, dp, ldd, 1,
                               int(accumulate), _stream()))


def copy_cl(src, dst):
    sp, lds, rows, c = cl_view(src)
    dp, ldd, rows2, c2 = cl_view(dst)
    assert rows == rows2 and c == c2
    check(lib().dcv_copy_cl(dcv_dtype(src), sp, lds, dcv_dtype(dst), dp, ldd, rows, c, _stream()))


# ------------------------------------------------------------------------------------ GRU, loss, Adam
def gru_traj_fwd(h0, eps, w_ih, w_hh, b_ih, b_hh):
    b, d = h0.shape
    t = eps.shape[0]
    hs = torch.empty((b, t, d), dtype=torch.float32, device=h0.device)
    check(lib().dcv_gru_traj_fwd(h0.data_ptr(), eps.data_ptr(), w_ih.data_ptr(), w_hh.data_ptr(), b_ih.data_ptr(),
                                 b_hh.data_ptr(), b, t, d, hs.data_ptr(), _stream()))
    return hs


def gru_traj_bwd(h0, eps, hs, dhs, w_ih, w_hh, b_ih, b_hh, dw_ih, dw_hh, db_ih, db_hh, accumulate=False):
    b, d = h0.shape
    t = eps.shape[0]
    assert dhs.is_contiguous() and dhs.dtype == torch.float32
    check(lib().dcv_gru_traj_bwd(h0.data_ptr(), eps.data_ptr(), hs.data_ptr(), dhs.data_ptr(), w_ih.data_ptr(),
                                 w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(), b, t, d, dw_ih.data_ptr(),
                                 dw_hh.data_ptr(), db_ih.data_ptr(), db_hh.data_ptr(), int(accumulate), _stream()))


def loss_fwd_bwd(y, kind, loss_out, accumulate, dy=None, grad_scale=1.0):
    """dense torch logits (module API)"""
    assert y.is_contiguous()
    check(lib().dcv_loss_fwd_bwd(dcv_dtype(y), y.data_ptr(), 1, y.numel(), kind, loss_out.data_ptr(), int(accumulate),
                                 _p(dy), 1, grad_scale, _stream()))


def loss_fwd_bwd_act(y, kind, loss_out, accumulate, dy=None, grad_scale=1.0):
    """single-channel logits Act (possibly with padding channels) - the fused trainer path"""
    assert y.c == 1 and (dy is None or dy.c == 1)
    check(lib().dcv_loss_fwd_bwd(dcv_dtype(y), y.ptr, y.ld, y.rows, kind, loss_out.data_ptr(), int(accumulate),
                                 None if dy is None else dy.ptr, 1 if dy is None else dy.ld, grad_scale, _stream()))


def adam_multi(params, grads, exp_avgs, exp_avg_sqs, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    n = len(params)
    arr = C.c_void_p * n
    numel = (C.c_int64 * n)(*[p.numel() for p in params])
    for group in (params, grads, exp_avgs, exp_avg_sqs):
        for t in group:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
    check(lib().dcv_adam_multi(n, arr(*[p.data_ptr() for p in params]), arr(*[g.data_ptr() for g in grads]),
                               arr(*[m.data_ptr() for m in exp_avgs]), arr(*[v.data_ptr() for v in exp_avg_sqs]), numel,
                               lr, beta1, beta2, eps, weight_decay, step, grad_scale, _stream()))
