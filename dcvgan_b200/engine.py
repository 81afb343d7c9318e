"""Forward / backward plans of the five DCVGAN networks over the C-ABI kernels.

A plan works on channels-last `Act` buffers, writes producers straight into channel slices of the
buffers their consumers read (no torch.cat), and keeps exactly what the backward pass needs.
The nn.Module classes in generator.py / discriminator.py own the parameters (same state_dict keys
as the reference) and delegate here; trainer.py calls the same plans without autograd.

Reference call stacks mirrored here: generator.py:84-141 (ggen), :361-435 (cgen),
discriminator.py:106-127, :210-231, :309-333 (idis, vdis, gdis).
"""
import os

import torch

from . import ops
from ._lib import ACT_LEAKY, ACT_NONE, ACT_TANH
from .ops import Act, ConvSpec

BN_EPS, BN_MOM = 1e-5, 0.1
# BatchNorm batch statistics accumulated in the convolution's own epilogue (dcv_conv_stats) for output tensors of at most this
# many MiB (0 = never; DCV_FUSED_BN_STATS=1 = always).  Measured per iteration at the end of round 2 (mug-depth, batch 32, one
# session): off 9.465 ms, <= 2 MiB 9.424, <= 8 9.414, <= 32 9.393, <= 70 9.362, <= 140 9.360, always 9.372 - every tensor but
# the 268 MiB output of cgen up_blocks.5 (whose small-K convolution is bound by its store epilogue) gains: the separate
# statistics pass costs a read of the tensor plus ~10 us of fixed latency per launch.
FUSED_BN_STATS = os.environ.get("DCV_FUSED_BN_STATS", "0") == "1"
FUSED_BN_STATS_MAX_MB = float(os.environ.get("DCV_FUSED_BN_STATS_MAX_MB", "140"))
TAP_UNROLL = os.environ.get("DCV_NO_TAP_UNROLL", "0") != "1"   # tiny-Cout transposed convolutions as 1x1 GEMM + col2im


# ------------------------------------------------------------------------------------------ RNG
class Rng:
    """Source of every random draw on the path (latents, Noise layers, Dropout2d masks).

    mode 'device'     : torch's CUDA generator (production / benchmark)
    mode 'cpu_parity' : the draw the reference would make on the CPU (torch.empty(shape).normal_() /
                        .bernoulli_()), in the reference's order, copied to the GPU - so that a run
                        seeded like the CPU oracle consumes bit-identical noise.
    mode 'staged'     : the same CPU draws, but delivered through STATIC device buffers that are filled before the
                        iteration starts (`stage()`), so that the iteration itself issues no host work and can be
                        captured into / replayed from a CUDA graph: parity tests of the graph path
                        (tests/test_timed_path_gpu.py).  The first iteration of a draw pattern records the request list.
    """

    def __init__(self, mode="device"):
        assert mode in ("device", "cpu_parity", "staged")
        self.mode = mode
        self.plan = []        # staged: [(kind, shape, device buffer)] in request order
        self.cursor = 0
        self.recording = True

    # ---- the reference's CPU draws
    @staticmethod
    def _cpu_normal(shape):
        return torch.empty(shape).normal_()

    @staticmethod
    def _cpu_noise(a):
        # normal_() fills in memory order, so (N,C,H,W) and (N,C,1,H,W) draws are identical
        return torch.empty((a[0], a[4], a[1], a[2], a[3])).normal_()

    @staticmethod
    def _cpu_dropout(n, c, p):
        return torch.empty((n, c, 1, 1)).bernoulli_(1 - p).div_(1 - p).view(n, c).contiguous()

    # ---- staged mode
    def stage(self):
        """draw every request of the recorded pattern on the CPU (reference order) into the static device buffers"""
        assert self.mode == "staged"
        self.cursor = 0
        if not self.plan:
            self.recording = True
            return
        self.recording = False
        for kind, arg, buf in self.plan:
            if kind == "normal":
                buf.copy_(self._cpu_normal(arg))
            elif kind == "noise":
                ref = self._cpu_noise(arg).cuda()
                ops.to_channels_last(ref, Act.from_dense(buf.view(arg)))
            else:
                buf.copy_(self._cpu_dropout(*arg))
        torch.cuda.synchronize()

    def _staged(self, kind, arg, make):
        if self.recording:
            buf = make()
            self.plan.append((kind, arg, buf))
            return buf
        k, a, buf = self.plan[self.cursor]
        assert (k, a) == (kind, arg), f"staged RNG: request {self.cursor} is {(kind, arg)}, recorded {(k, a)}"
        self.cursor += 1
        return buf

    def normal(self, shape):
        if self.mode == "cpu_parity":
            return self._cpu_normal(shape).cuda()
        if self.mode == "staged":
            return self._staged("normal", tuple(shape), lambda: self._cpu_normal(shape).cuda())
        return torch.empty(shape, device="cuda").normal_()

    def noise_for(self, a):
        """fp32 noise laid out like `a` (dense channels-last); drawn in the reference's (N,C[,T],H,W) order"""
        if self.mode in ("cpu_parity", "staged"):
            def make():
                ref = self._cpu_noise(a.shape).cuda()
                out = Act.empty(a.n, a.t, a.h, a.w, a.c, torch.float32)
                ops.to_channels_last(ref, out)
                return out.base
            if self.mode == "staged":
                return self._staged("noise", tuple(a.shape), make)
            return make()
        return torch.empty(a.rows * a.c, device="cuda").normal_()

    def dropout_scale(self, n, c, p=0.5):
        """Dropout2d: one Bernoulli(1-p) per (sample, channel), scaled by 1/(1-p)  (generator.py:246-248)"""
        if self.mode == "cpu_parity":
            return self._cpu_dropout(n, c, p).cuda()
        if self.mode == "staged":
            return self._staged("dropout", (n, c, p), lambda: self._cpu_dropout(n, c, p).cuda())
        return torch.empty((n, c), device="cuda").bernoulli_(1 - p).div_(1 - p)


_RNG = Rng("device")


def set_rng_mode(mode):
    global _RNG
    _RNG = Rng(mode)


def rng():
    return _RNG


# ------------------------------------------------------------------------------------------ gradient sinks
class GradSink:
    """Where parameter gradients go.  `get(p)` returns (fp32 tensor shaped like p, accumulate?)."""

    def __init__(self, lookup=None):
        self.lookup = lookup          # optional dict: param -> preallocated grad tensor
        self.grads = {}

    def get(self, p):
        if p in self.grads:
            return self.grads[p], True
        g = self.lookup[p] if self.lookup is not None else torch.empty_like(p, memory_format=torch.contiguous_format)
        self.grads[p] = g
        return g, False


# ------------------------------------------------------------------------------------------ packed-weight cache
# The fused trainer re-uses packed (bf16, K-major) weights between the passes that see the same master weights
# (D: real + fake in the D-phase; G: forward + backward).  WCACHE maps id(param) -> {(geom, dir, impl): packed}; the
# trainer drops a network's entries right after its Adam step.  None (module / autograd API) = always re-pack.
WCACHE = None


# Batch packing: PACK_GROUP maps id(weight) -> the group (network) it belongs to, PACK_LOG[group] remembers every
# (spec, geometry, direction, weight) that was ever packed for the group.  A cache miss on one weight then re-packs the
# WHOLE group in one launch (ops.pack_weight_batch) - after a network's Adam step all its packs are stale together.
PACK_GROUP = None
PACK_LOG = None     # owned by the trainer like WCACHE (group -> {key: job})


def packed_weight(spec, g, direction, impl, weight):
    if WCACHE is None:
        return ops.pack_weight(spec, g, direction, impl, weight)
    per = WCACHE.setdefault(id(weight), {})
    key = (g.key(), direction, impl)
    wp = per.get(key)
    if wp is not None:
        return wp
    group = PACK_GROUP.get(id(weight)) if (PACK_GROUP is not None and impl == ops.IMPL_TC) else None
    if group is None:
        wp = per[key] = ops.pack_weight(spec, g, direction, impl, weight)
        return wp
    log = PACK_LOG.setdefault(group, {})
    log.setdefault((id(weight),) + key, (spec, g, direction, weight))
    jobs = [(k, j) for k, j in log.items() if (k[1:]) not in WCACHE.get(k[0], {})]
    outs = ops.pack_weight_batch([j for _, j in jobs])
    for (k, _), o in zip(jobs, outs):
        WCACHE.setdefault(k[0], {})[k[1:]] = o
    return per[key]


def invalidate_packed(params):
    if WCACHE is not None:
        for p in params:
            WCACHE.pop(id(p), None)


DEFER_BN = os.environ.get("DCV_NO_DEFER_BN", "0") != "1"    # cgen up_blocks.5 without backward: BatchNorm + ReLU applied by Outconv on load


# ------------------------------------------------------------------------------------------ one layer group
class Block:
    """[Noise] -> conv -> [BatchNorm] -> [Dropout2d] -> activation, with its backward.

    conv / bn are the owner's real nn.Conv*/nn.BatchNorm* modules (parameter containers only).
    """

    def __init__(self, spec, conv, bn=None, act=ACT_NONE, slope=0.0, dropout=False, noise=None):
        self.spec, self.conv, self.bn = spec, conv, bn
        self.act, self.slope, self.dropout = act, slope, dropout
        self.noise = noise  # None or (use_noise, sigma)

    def img_scatter_paths_ok(self, x, out):
        """True when the forward of this (BatchNorm-free, transposed) layer runs on the image-side scatter kernel, which can apply
        a producer's deferred BatchNorm while it loads x"""
        spec = self.spec
        if self.bn is not None or spec.kind != "convT" or (self.noise is not None and self.noise[0]):
            return False
        g = spec.geom(x.n, x.spatial, x.cp, out.cp)
        return ops.img_conv_ok(spec, g, ops.IMG_SCATTER, out, x)

    def forward(self, x, out, training, rng_, save=True, loss=None, defer_bn=False, pre=None):
        """x: input Act; out: Act (slice) that receives the block's output.  Returns ctx for backward.
        loss (heads only): {'kind', 'out' (1-element fp32 tensor), 'accumulate', 'want_grad'} - the adversarial-loss term of
        this head is computed in the head's own launch; ctx['dlogits'] then holds dL/dlogits (or None).
        defer_bn (forward passes that are not differentiated): leave the PRE-BatchNorm convolution output z in `out` and return
        the normalisation in ctx['pre'] - the only consumer (Outconv) applies BatchNorm + ReLU while it loads the tensor, so the
        normalise pass over it never runs.  pre: such a deferred normalisation of channels [c0, c0 + 64) of x."""
        spec = self.spec
        w = self.conv.weight
        x_used = x
        if self.noise is not None and self.noise[0]:
            x_used = x.like()
            ops.add_noise(x, rng_.noise_for(x), float(self.noise[1]), x_used)
        z = out if (self.bn is None or defer_bn) else Act.empty(out.n, out.t, out.h, out.w, out.c, out.dtype)
        # geometry over the zero-padded channel counts of both buffers (dcv_geom.wCl/wCs carry the real ones)
        cin_p, cout_p = x_used.cp, z.cp
        g = spec.geom(x.n, x.spatial, cin_p, cout_p)
        ctx = {"g": g, "x": x_used if save else None, "a": out, "cin_p": cin_p, "cout_p": cout_p}
        assert pre is None or not save, "a deferred BatchNorm is for forward passes that are not differentiated"
        if (self.bn is None and spec.kind == "conv" and self.act in (ACT_NONE, ACT_LEAKY) and ops.img_conv_ok(spec, g, ops.IMG_FWD, x_used, out)
                and ops.img_conv_ok(spec, g, ops.IMG_BWD, x_used, out)):
            ops.img_conv_fwd(spec, g, x_used, w, out, self.act, self.slope)       # Inconv: HBM-bound mma.sync kernel
            ctx["img"] = True
            return ctx
        if self.bn is None and spec.kind == "convT" and ops.img_conv_ok(spec, g, ops.IMG_SCATTER, out, x_used):
            ops.img_conv_scatter(spec, g, x_used, w, out, self.act, self.slope, pre)    # Outconv: one pass over the 128-channel input
            return ctx
        assert pre is None, "a deferred BatchNorm needs the image-side kernels of its consumer"
        impl = ops.choose_conv_impl(g, spec.fwd_dir, x_used)
        wp = packed_weight(spec, g, spec.fwd_dir, impl, w)
        if self.bn is None:
            if self._tap_unrolled(impl, x_used):
                self._forward_tap_unrolled(x_used, out)
                return ctx
            if loss is not None and self.act == ACT_NONE and spec.cout == 1 and ops.head_loss_ok(g, spec.fwd_dir, impl, x_used):
                ctx["dlogits"] = ops.head_loss(g, spec.fwd_dir, x_used.padded_to(cin_p), wp, out.padded_to(cout_p), loss["kind"],
                                               loss["out"], loss["accumulate"], loss["want_grad"])
                ctx["loss_fused"] = True
                return ctx
            ops.conv(g, spec.fwd_dir, impl, x_used.padded_to(cin_p), wp, out.padded_to(cout_p), self.act, self.slope)
            return ctx
        bn = self.bn
        zp = z.padded_to(cout_p)
        # statistics in the convolution's epilogue where that pays (see FUSED_BN_STATS_MAX_MB), else the streaming pass
        fuse_stats = FUSED_BN_STATS or z.rows * z.c * 2 <= FUSED_BN_STATS_MAX_MB * (1 << 20)
        slots = ops.conv_stats_slots(g, spec.fwd_dir, x_used, zp) if (fuse_stats and training and impl == ops.IMPL_TC and zp.c == z.c) else 0
        if slots > 0:
            partials = ops.conv_stats(g, spec.fwd_dir, x_used.padded_to(cin_p), wp, zp, slots)
            mean, invstd = ops.bn_finalize(partials, z.rows, BN_EPS, BN_MOM, bn.running_mean, bn.running_var, bn.num_batches_tracked)
        else:
            ops.conv(g, spec.fwd_dir, impl, x_used.padded_to(cin_p), wp, zp)
            if training:
                mean, invstd = ops.bn_batch_stats(z, BN_EPS, BN_MOM, bn.running_mean, bn.running_var, bn.num_batches_tracked)
            else:
                mean, invstd = ops.bn_eval_stats(bn.running_mean, bn.running_var, BN_EPS)
        if defer_bn:
            assert not self.dropout and self.act == ACT_LEAKY
            ctx["pre"] = {"mean": mean, "invstd": invstd, "gamma": bn.weight.detach(), "beta": bn.bias.detach(), "slope": self.slope}
            return ctx
        drop = rng_.dropout_scale(z.n, z.c) if (self.dropout and training) else None
        ops.bn_act(z, mean, invstd, bn.weight.detach(), bn.bias.detach(), drop, self.act, self.slope, out)
        if save:
            ctx.update(z=z, mean=mean, invstd=invstd, drop=drop, training=training)
        return ctx

    # ---- transposed convolutions with <= 4 output channels (outconv 128 -> 3, ggen main.12 64 -> C): 1x1 GEMM + col2im
    def _tap_unrolled(self, impl, x):
        spec = self.spec
        return (TAP_UNROLL and impl == ops.IMPL_TC and spec.kind == "convT" and spec.k[0] == 1 and x.t == 1 and spec.cout <= 4
                and spec.cout * spec.taps <= 64 and spec.cin % 16 == 0 and spec.s[1] == spec.s[2] and spec.p[1] == spec.p[2])

    def _forward_tap_unrolled(self, x, out):
        """The forward of such a layer as a tensor-core tile has N = 16 padded columns and re-reads a 128x16 A slice from
        shared memory for every tap (outconv: 0.239 ms at B = 32, bound by shared-memory operand reads).  Here the input
        is read once by a 1x1 convolution onto Cout*taps columns (the ConvTranspose2d weight (Cin, Cout, kh, kw) already
        is that matrix) and a col2im pass sums the taps per output pixel and applies the activation.  The backward pass
        is unchanged (data / weight gradients of the original geometry)."""
        spec = self.spec
        ncol = spec.cout * spec.taps
        spec1 = ConvSpec("conv", spec.cin, ncol, (1, 1, 1), (1, 1, 1), (0, 0, 0))
        P = Act.empty(x.n, 1, x.h, x.w, ncol, x.dtype)
        g1 = spec1.geom(x.n, x.spatial, x.cp, P.cp)
        impl1 = ops.choose_conv_impl(g1, spec1.fwd_dir, x)
        w = self.conv.weight
        key = ("unrolled", g1.key(), impl1)
        per = WCACHE.setdefault(id(w), {}) if WCACHE is not None else {}
        wp = per.get(key)
        if wp is None:    # w[ci][co][tap] seen as w[cl = ci][cs = co*taps + tap]: strides (ncol, 1), a single "tap"
            wp = per[key] = ops.pack_weight_strided(g1, spec1.fwd_dir, impl1, w, ncol, 1, 0)
        ops.conv(g1, spec1.fwd_dir, impl1, x.padded_to(x.cp), wp, P.padded_to(P.cp))
        ops.col2im_act(P, out, spec.cout, spec.k[1], spec.k[2], spec.s[1], spec.p[1], self.act, self.slope)

    def backward(self, ctx, da, sink, dx_out=None, need_dw=True, dx_accumulate=False):
        """da: gradient w.r.t. the block output (Act, may be a slice).  dx_out: Act receiving dL/dx or None.
        dx_accumulate: dx_out already holds a gradient (a U-Net skip connection's) and dL/dx is ADDED to it - by the
        convolution's own epilogue (TMA reduce-add) when the tcgen05 path allows, else through a temporary and an add."""
        spec, g = self.spec, ctx["g"]
        a = ctx["a"]
        if ctx.get("img"):       # activation derivative, weight gradient and data gradient in one pass over (da, a)
            dw, acc = sink.get(self.conv.weight) if need_dw else (None, False)
            ops.img_conv_bwd(spec, g, da, a, ctx["x"], self.conv.weight, self.act, self.slope, dw, acc, dx_out)
            return None
        if self.bn is not None:
            dz = ctx["z"].like()
            if need_dw:
                dgam, acc = sink.get(self.bn.weight)
                dbet, _ = sink.get(self.bn.bias)
            else:
                dgam = dbet = None
                acc = False
            ops.bn_act_bwd(da, a, ctx["z"], ctx["mean"], ctx["invstd"], self.bn.weight.detach(), self.bn.bias.detach(), ctx["drop"],
                           self.act, self.slope, dz, dgam, dbet, acc, batch_stats=ctx.get("training", True))
        elif self.act != ACT_NONE:
            dz = Act.empty(a.n, a.t, a.h, a.w, a.c, a.dtype)
            ops.act_bwd(da, a, self.act, self.slope, dz)
        else:
            dz = da
        # dz / dx_out are viewed with the same zero-padded widths the forward geometry was built with
        cin_p, cout_p = ctx["cin_p"], ctx["cout_p"]
        dzp = dz.padded_to(cout_p)
        # Outconv (ConvTranspose2d(128, 3, 3, 1, 1), generator.py:272-277): its 3-channel side is the L tensor of the geometry -
        # the weight gradient streams the 128-channel input once and the data gradient is a forward-type pass 3 -> 128
        # channels, both through the image-side mma.sync kernels instead of 16-channel-padded tensor-core tiles
        img = spec.kind == "convT" and self.bn is None
        if need_dw:
            dw, acc = sink.get(self.conv.weight)
            xp = ctx["x"].padded_to(cin_p)
            xl, xs = (xp, dzp) if spec.kind == "conv" else (dzp, xp)
            if img and ops.img_conv_ok(spec, g, ops.IMG_WGRAD, dz, ctx["x"]):
                ops.img_conv_bwd(spec, g, ctx["x"], None, dz, self.conv.weight, ACT_NONE, 0.0, dw, acc, None)
            else:
                ops.wgrad(spec, g, xl, xs, dw, accumulate=acc)
        if dx_out is not None:
            if img and ops.img_conv_ok(spec, g, ops.IMG_FWD, dz, dx_out):
                ops.img_conv_fwd(spec, g, dz, self.conv.weight, dx_out, ACT_NONE, 0.0)
            else:
                impl = ops.choose_conv_impl(g, spec.bwd_dir, dzp)
                wp = packed_weight(spec, g, spec.bwd_dir, impl, self.conv.weight)
                dxp = dx_out.padded_to(cin_p)
                if not dx_accumulate:
                    ops.conv(g, spec.bwd_dir, impl, dzp, wp, dxp)
                elif ops.conv_accumulate_ok(g, spec.bwd_dir, impl, dzp, dxp):
                    ops.conv_accumulate(g, spec.bwd_dir, dzp, wp, dxp)
                else:
                    tmp = Act.empty(dx_out.n, dx_out.t, dx_out.h, dx_out.w, dx_out.c, dx_out.dtype)
                    ops.conv(g, spec.bwd_dir, impl, dzp, wp, tmp.padded_to(cin_p))
                    ops.axpy(tmp, dx_out, True)
        return dz


class MergedStem:
    """The two stem convolutions of idis / vdis (conv_g on the geometry channels, conv_c on the colour channels, both
    + LeakyReLU(0.2), outputs concatenated as [hc | hg]: discriminator.py:79-90,121-124,180-193,225-228) run as ONE
    convolution over the concatenated input [xg | xc] with a block-structured weight - conv_g occupies input channels
    [0,cg) x output channels [ndf/2, ndf), conv_c input channels [cg, cg+cc) x output channels [0, ndf/2), zeros
    elsewhere.  Same arithmetic (the zero blocks contribute exact zeros to the fp32 accumulators); one 16/32-channel
    operand instead of two zero-padded ones, a 2x wider N tile, and channel halves that need not be multiples of 16
    (vdis.ndf 48 in config/surreal-segm.yml).  bf16 / tcgen05 path only.

    With few input channels (4 * (cg + cc) <= 32: depth, optical flow) the four taps along w are additionally folded
    into the channel dimension (ops.fold_w): the convolution proper then has kw = 1 and 4*(cg+cc) real channels per
    tap - a quarter of the TMA boxes and MMAs of the 16-channel-pitch form, which was bound by the TMA request rate
    (vdis stem, B = 32: forward 0.133 -> 0.048 ms, weight gradient 0.134 -> 0.054, data gradient 0.218 -> 0.074)."""

    KW, SW, PW = 4, 2, 1      # stem kernels are 4 wide, stride 2, padding 1 along w (both discriminators)

    def __init__(self, kind, cg, cc, ndf, conv_g, conv_c, noise=None):
        self.cg, self.cc, self.ndf = cg, cc, ndf
        self.conv_g, self.conv_c, self.noise = conv_g, conv_c, noise
        cin, half, kt = cg + cc, ndf // 2, (4 if kind == "vdis" else 1)
        self.fold = FOLD_STEMS and self.KW * cin <= 32
        if self.fold:
            self.spec = ConvSpec("conv", self.KW * cin, ndf, (kt, 4, 1), (1, 2, 1), (0, 1, 0))
            self.windows = [(m, k * cin + c0, cs0, k) for k in range(self.KW) for m, c0, cs0 in ((conv_g, 0, half), (conv_c, cg, 0))]
        else:
            self.spec = conv3d_spec(cin, ndf) if kind == "vdis" else conv2d_spec(cin, ndf, 4, 2, 1)
            self.windows = [(conv_g, 0, half, None), (conv_c, cg, 0, None)]      # (module, cl_off, cs_off, folded w tap)

    def _part(self, tensor, cl_off, cs_off, k):
        return ops.conv_weight_part(tensor, cl_off, cs_off, k, self.KW)

    def _packed(self, g, direction, impl):
        parts = [self._part(m.weight, cl, cs, k) for m, cl, cs, k in self.windows]
        if WCACHE is None:
            return ops.pack_weight_merged(g, direction, impl, parts)
        per = WCACHE.setdefault(id(self.conv_g.weight), {})
        key = ("merged", g.key(), direction, impl)
        wp = per.get(key)
        if wp is None:
            wp = per[key] = ops.pack_weight_merged(g, direction, impl, parts)
        return wp

    def forward(self, xg, xc, out, training, rng_, save=True):
        cg, cin = self.cg, self.cg + self.cc
        noisy = self.noise is not None and self.noise[0]
        sigma = float(self.noise[1]) if noisy else 0.0
        ng = rng_.noise_for(xg) if noisy else None                               # draw order: geometry, then colour
        nc = rng_.noise_for(xc) if noisy else None
        if self.fold:
            ow = (xg.w + 2 * self.PW - self.KW) // self.SW + 1
            cat = Act.empty(xg.n, xg.t, xg.h, ow, self.KW * cin, xg.dtype)
            ops.fold_w(xg, xc, cat, self.KW, self.SW, self.PW, ng, nc, sigma)
        else:
            cat = Act.empty(xg.n, xg.t, xg.h, xg.w, cin, xg.dtype)
            if noisy:
                ops.add_noise(xg, ng, sigma, cat.ch(0, cg))
                ops.add_noise(xc, nc, sigma, cat.ch(cg, cin))
            else:
                ops.copy_cl(xg, cat.ch(0, cg))
                ops.copy_cl(xc, cat.ch(cg, cin))
        g = self.spec.geom(cat.n, cat.spatial, cat.cp, out.cp)
        impl = ops.choose_conv_impl(g, self.spec.fwd_dir, cat)
        ops.conv(g, self.spec.fwd_dir, impl, cat.padded_to(cat.cp), self._packed(g, self.spec.fwd_dir, impl),
                 out.padded_to(out.cp), ACT_LEAKY, 0.2)
        return {"g": g, "x": cat if save else None, "a": out, "in_shape": xg.shape}

    def backward(self, ctx, da, sink, need_dx, need_dw):
        """Returns (dxg, dxc) or (None, None)."""
        g, a, cat = ctx["g"], ctx["a"], ctx["x"]
        dz = Act.empty(a.n, a.t, a.h, a.w, a.c, a.dtype)
        ops.act_bwd(da, a, ACT_LEAKY, 0.2, dz)
        dzp = dz.padded_to(dz.cp)
        if need_dw:
            parts, got = [], {}
            for m, cl, cs, k in self.windows:
                if m not in got:                       # the windows of one weight are disjoint: same accumulate flag for all
                    got[m] = sink.get(m.weight)
                dw, acc = got[m]
                parts.append((self._part(dw, cl, cs, k), acc))
            ops.wgrad_merged(g, cat.padded_to(cat.cp), dzp, parts)
        if not need_dx:
            return None, None
        dcat = Act.empty(cat.n, cat.t, cat.h, cat.w, cat.c, cat.dtype)
        impl = ops.choose_conv_impl(g, self.spec.bwd_dir, dzp)
        ops.conv(g, self.spec.bwd_dir, impl, dzp, self._packed(g, self.spec.bwd_dir, impl), dcat.padded_to(dcat.cp))
        if not self.fold:
            return dcat.ch(0, self.cg), dcat.ch(self.cg, self.cg + self.cc)
        n, t, h, w, _ = ctx["in_shape"]
        dxg = Act.empty(n, t, h, w, self.cg, cat.dtype)
        dxc = Act.empty(n, t, h, w, self.cc, cat.dtype)
        ops.unfold_w(dcat, dxg, dxc, self.KW, self.SW, self.PW)
        return dxg, dxc


FOLD_STEMS = True    # tests flip it to compare the folded form against the plain merged convolution
MERGE_STEMS = True   # bf16 path; tests flip it to compare against the two-convolution form


def _k2(k):
    return (1, k, k)


def conv2d_spec(cin, cout, k, s, p):
    return ConvSpec("conv", cin, cout, _k2(k), _k2(s), (0, p, p))


def convT2d_spec(cin, cout, k, s, p):
    return ConvSpec("convT", cin, cout, _k2(k), _k2(s), (0, p, p))


def conv3d_spec(cin, cout):
    return ConvSpec("conv", cin, cout, (4, 4, 4), (1, 2, 2), (0, 1, 1))


# ------------------------------------------------------------------------------------------ ggen
class GGenPlan:
    """GRU latent trajectory + 5 transposed convolutions (generator.py:58-80,84-141)."""

    def __init__(self, mod):
        self.mod = mod
        m = mod.main
        ngf, dz, C = mod.ngf, mod.dim_z, mod.channel
        chans = [dz, ngf * 8, ngf * 4, ngf * 2, ngf, C]
        self.blocks = []
        for i in range(5):
            spec = convT2d_spec(chans[i], chans[i + 1], 4, 1 if i == 0 else 2, 0 if i == 0 else 1)
            if i < 4:
                self.blocks.append(Block(spec, m[3 * i], m[3 * i + 1], ACT_LEAKY, 0.0))       # BN + ReLU
            else:
                self.blocks.append(Block(spec, m[12], None, ACT_NONE if mod.geometric_info == "segmentation" else ACT_TANH))
        self.softmax = mod.geometric_info == "segmentation"

    def forward(self, B, training, dtype, rng_, save=True):
        mod = self.mod
        T, dzc, dzm = mod.video_length, mod.dim_z_content, mod.dim_z_motion
        rec = mod.recurrent
        z_c = rng_.normal((B, dzc))                                              # draw order: generator.py:103-116
        h0 = rng_.normal((B, dzm))
        if rng_.mode == "device":      # one draw for the whole trajectory (the per-step draws of generator.py:84-101 are i.i.d.)
            eps = rng_.normal((T, B, dzm))
        else:                          # CPU-parity runs consume the reference's draws one by one
            eps = torch.stack([rng_.normal((B, dzm)) for _ in range(T)], 0).contiguous()
        P = [p.detach() for p in (rec.weight_ih, rec.weight_hh, rec.bias_ih, rec.bias_hh)]
        hs = ops.gru_traj_fwd(h0, eps, *P)                                       # (B, T, dzm)
        x = Act.empty(B * T, 1, 1, 1, dzc + dzm, dtype)
        z = x.rows2d().view(B, T, dzc + dzm) if x.ld == dzc + dzm else x.base.as_strided((B, T, dzc + dzm), (T * x.ld, x.ld, 1))
        z[:, :, :dzc] = z_c[:, None, :]                                          # z_content repeated over T (:103-108)
        z[:, :, dzc:] = hs
        ctxs = []
        sp = (1, 1, 1)
        for blk in self.blocks:
            sp = blk.spec.out_spatial(sp)
            out = Act.empty(B * T, 1, sp[1], sp[2], blk.spec.cout, dtype)
            ctxs.append(blk.forward(x, out, training, rng_, save))
            x = out
        if self.softmax:
            y = x.like()
            ops.softmax(x, y)
            x = y
        ctx = {"blocks": ctxs, "h0": h0, "eps": eps, "hs": hs, "out": x, "B": B} if save else None
        return x, ctx

    def backward(self, ctx, dout, sink):
        """dout: Act (B*T,1,64,64,C) gradient w.r.t. the generated frames."""
        mod = self.mod
        da = dout
        if self.softmax:
            dz = dout.like()
            ops.softmax_bwd(dout, ctx["out"], dz)
            da = dz
        for i in range(4, -1, -1):
            blk, c = self.blocks[i], ctx["blocks"][i]
            x = c["x"]
            dx = Act.empty(x.n, x.t, x.h, x.w, x.c, x.dtype)
            blk.backward(c, da, sink, dx_out=dx)
            da = dx
        B, T, dzc, dzm = ctx["B"], mod.video_length, mod.dim_z_content, mod.dim_z_motion
        dhs = da.base.as_strided((B, T, dzc + dzm), (T * da.ld, da.ld, 1))[:, :, dzc:].float().contiguous()
        rec = mod.recurrent
        P = [p.detach() for p in (rec.weight_ih, rec.weight_hh, rec.bias_ih, rec.bias_hh)]
        grads = [sink.get(p) for p in (rec.weight_ih, rec.weight_hh, rec.bias_ih, rec.bias_hh)]
        ops.gru_traj_bwd(ctx["h0"], ctx["eps"], ctx["hs"], dhs, *P, *[g for g, _ in grads], accumulate=grads[0][1])


# ------------------------------------------------------------------------------------------ cgen
class CGenPlan:
    """U-Net colour generator (generator.py:158-282,323-402)."""

    def __init__(self, mod):
        self.mod = mod
        ngf, dz, C = mod.ngf, mod.dim_z, mod.in_ch
        self.ngf, self.dz, self.C = ngf, dz, C
        self.inconv = Block(conv2d_spec(C, ngf, 3, 1, 1), mod.inconv.main[0], None, ACT_LEAKY, 0.01)
        dch = [(ngf, ngf), (ngf, 2 * ngf), (2 * ngf, 4 * ngf), (4 * ngf, 4 * ngf), (4 * ngf, 4 * ngf), (4 * ngf, 4 * ngf)]
        self.down = [Block(conv2d_spec(ci, co, 4, 2, 1), mod.down_blocks[i].main[0], mod.down_blocks[i].main[1], ACT_LEAKY, 0.2)
                     for i, (ci, co) in enumerate(dch)]
        uch = [(4 * ngf + dz, 4 * ngf), (8 * ngf, 4 * ngf), (8 * ngf, 4 * ngf), (8 * ngf, 2 * ngf), (4 * ngf, ngf), (2 * ngf, ngf)]
        self.up = [Block(convT2d_spec(ci, co, 4, 2, 1), mod.up_blocks[i].main[0], mod.up_blocks[i].main[1], ACT_LEAKY, 0.0,
                         dropout=(i < 2)) for i, (ci, co) in enumerate(uch)]
        self.outconv = Block(convT2d_spec(2 * ngf, 3, 3, 1, 1), mod.outconv.main[0], None, ACT_TANH)
        # concat buffers: (spatial, [upsampled channels, skip channels])
        self.cat_layout = [(1, 4 * ngf, dz), (2, 4 * ngf, 4 * ngf), (4, 4 * ngf, 4 * ngf), (8, 4 * ngf, 4 * ngf),
                           (16, 2 * ngf, 2 * ngf), (32, ngf, ngf), (64, ngf, ngf)]

    def forward(self, x, z, training, rng_, save=True):
        """x: Act (N,1,64,64,C) geometry frames; z: fp32 (N, dim_z) colour code per frame."""
        mod, N, dtype = self.mod, x.n, x.dtype
        if mod.geometric_info == "segmentation":                                 # generator.py:378-385
            xr = x.like()
            ops.segm_remap(x, xr)
            x = xr
        cats = [Act.empty(N, 1, s, s, a + b, dtype) for (s, a, b) in self.cat_layout]
        # cats[0] = [down5 | z] @1, cats[k] = [up_{k-1} | skip] for k=1..5, cats[6] = [up5 | inconv] @64
        first = [a for (_, a, _) in self.cat_layout]
        ctx = {"cats": cats}
        ctx["inconv"] = self.inconv.forward(x, cats[6].ch(first[6], cats[6].c), training, rng_, save)
        src = cats[6].ch(first[6], cats[6].c)
        dctx = []
        for i in range(6):
            dst = cats[5 - i].ch(first[5 - i], cats[5 - i].c) if i < 5 else cats[0].ch(0, first[0])
            dctx.append(self.down[i].forward(src, dst, training, rng_, save))
            src = dst
        zc = cats[0].ch(first[0], cats[0].c)                                     # generator.py:393
        zc.rows2d().copy_(z)
        uctx = []
        out = Act.empty(N, 1, 64, 64, 3, dtype)
        # up_blocks.5's BatchNorm + ReLU output is read by Outconv only: when no backward pass follows (the fake batch of the
        # D-phase, sampling) leave the convolution output in the concat buffer and let Outconv's kernel normalise on load (no
        # bn_act pass over the 64x64x64 tensor; bit-identical operands)
        defer = (DEFER_BN and not save and dtype == torch.bfloat16 and first[6] == 64
                 and self.outconv.img_scatter_paths_ok(cats[6], out))
        for i in range(6):
            dst = cats[i + 1].ch(0, first[i + 1])
            uctx.append(self.up[i].forward(cats[i], dst, training, rng_, save, defer_bn=(defer and i == 5)))
        pre = dict(uctx[5]["pre"], c0=0) if defer else None
        ctx["outconv"] = self.outconv.forward(cats[6], out, training, rng_, save, pre=pre)
        ctx.update(down=dctx, up=uctx)
        return out, (ctx if save else None)

    def backward(self, ctx, dout, sink, need_dx=True):
        cats = ctx["cats"]
        first = [a for (_, a, _) in self.cat_layout]
        dcats = [c.like() for c in cats]
        self.outconv.backward(ctx["outconv"], dout, sink, dx_out=dcats[6])
        for i in range(5, -1, -1):
            self.up[i].backward(ctx["up"][i], dcats[i + 1].ch(0, first[i + 1]), sink, dx_out=dcats[i])
        da = dcats[0].ch(0, first[0])
        for i in range(5, -1, -1):
            # the input of down block i is the skip tensor in the second half of concat buffer k: its gradient from the up path
            # is already there, and the block's data gradient is added on top of it (generator.py:398-400) - by the
            # convolution epilogue itself (TMA reduce-add), no separate a + b pass over the tensor
            k = 5 - i + 1 if i > 0 else 6
            skip = dcats[k].ch(first[k], dcats[k].c)
            self.down[i].backward(ctx["down"][i], da, sink, dx_out=skip, dx_accumulate=True)
            da = skip
        dx = None
        if need_dx and self.mod.geometric_info != "segmentation":                # argmax/scatter blocks the gradient
            xin = ctx["inconv"]["x"]
            dx = Act.empty(xin.n, xin.t, xin.h, xin.w, xin.c, xin.dtype)
        self.inconv.backward(ctx["inconv"], da, sink, dx_out=dx)
        return dx


# ------------------------------------------------------------------------------------------ discriminators
class DisPlan:
    """Image / video / gradient discriminator (discriminator.py:42-346)."""

    def __init__(self, mod, kind):
        self.mod, self.kind = mod, kind
        ndf, cg, cc = mod.ndf, mod.ch1, mod.ch2
        nz = (mod.use_noise, mod.noise_sigma)
        m = mod.main
        if kind == "idis":
            mk = lambda ci, co: conv2d_spec(ci, co, 4, 2, 1)
            self.stem_g = Block(mk(cg, ndf // 2), mod.conv_g[1], None, ACT_LEAKY, 0.2, noise=nz)
            self.stem_c = Block(mk(cc, ndf // 2), mod.conv_c[1], None, ACT_LEAKY, 0.2, noise=nz)
        elif kind == "vdis":
            mk = conv3d_spec
            self.stem_g = Block(mk(cg, ndf // 2), mod.conv_g[0], None, ACT_LEAKY, 0.2)
            self.stem_c = Block(mk(cc, ndf // 2), mod.conv_c[0], None, ACT_LEAKY, 0.2)
        else:
            mk = conv3d_spec
        self.merged = None
        if kind in ("idis", "vdis"):
            self.merged = MergedStem(kind, cg, cc, ndf, self.stem_g.conv, self.stem_c.conv, self.stem_g.noise)
        if kind in ("idis", "vdis"):
            self.main = [Block(mk(ndf, ndf * 2), m[1], m[2], ACT_LEAKY, 0.2, noise=nz),
                         Block(mk(ndf * 2, ndf * 4), m[5], m[6], ACT_LEAKY, 0.2, noise=nz),
                         Block(mk(ndf * 4, 1), m[9], None, ACT_NONE, noise=nz)]
        else:
            self.main = [Block(mk(cg, ndf), m[1], m[2], ACT_LEAKY, 0.2, noise=nz),
                         Block(mk(ndf, ndf * 2), m[5], m[6], ACT_LEAKY, 0.2, noise=nz),
                         Block(mk(ndf * 2, ndf * 4), m[9], m[10], ACT_LEAKY, 0.2, noise=nz),
                         Block(mk(ndf * 4, 1), m[13], None, ACT_NONE, noise=nz)]

    def forward(self, xg, xc, training, rng_, save=True, loss=None):
        """xg, xc: Acts (B,T,H,W,C) (T == 1 for idis).  Returns (logits Act (B, To, 4, 4, 1), ctx); with `loss` (see
        Block.forward) the head fuses its loss term and ctx_or_info['head'] carries {'loss_fused', 'dlogits'}."""
        ndf, dtype = self.mod.ndf, xg.dtype
        ctx = {}
        if self.kind == "gdis":
            h = Act.empty(xg.n, xg.t - 1, xg.h, xg.w, xg.c, dtype)               # discriminator.py:330-331
            ops.tdiff(xg, h)
            ctx["xg_shape"] = xg.shape
        else:
            sp = self.stem_g.spec.out_spatial(xg.spatial)
            h = Act.empty(xg.n, sp[0], sp[1], sp[2], ndf, dtype)
            # Noise draw order: geometry stem first, then colour (discriminator.py:121-122); cat = [hc, hg] (:124)
            if MERGE_STEMS and dtype == torch.bfloat16:
                ctx["stem"] = self.merged.forward(xg, xc, h, training, rng_, save)
            else:
                ctx["stem_g"] = self.stem_g.forward(xg, h.ch(ndf // 2, ndf), training, rng_, save)
                ctx["stem_c"] = self.stem_c.forward(xc, h.ch(0, ndf // 2), training, rng_, save)
        mctx = []
        for blk in self.main:
            sp = blk.spec.out_spatial(h.spatial)
            out = Act.empty(h.n, sp[0], sp[1], sp[2], blk.spec.cout, dtype)
            mctx.append(blk.forward(h, out, training, rng_, save, loss=loss if blk is self.main[-1] else None))
            h = out
        ctx["main"] = mctx
        self.last_head = {"loss_fused": bool(mctx[-1].get("loss_fused")), "dlogits": mctx[-1].get("dlogits")}
        return h, (ctx if save else None)

    def backward(self, ctx, dlogits, sink, need_dx=False, need_dw=True):
        """Returns (dxg, dxc) Acts or (None, None).  dxc is None for gdis (it ignores xc)."""
        da = dlogits
        n_main = len(self.main)
        for i in range(n_main - 1, -1, -1):
            x = ctx["main"][i]["x"]
            first = i == 0
            if first and self.kind == "gdis" and not need_dx:
                dx = None
            else:
                dx = Act.empty(x.n, x.t, x.h, x.w, x.c, x.dtype)
            self.main[i].backward(ctx["main"][i], da, sink, dx_out=dx, need_dw=need_dw)
            da = dx
        if self.kind == "gdis":
            if not need_dx:
                return None, None
            n, t, h, w, c = ctx["xg_shape"]
            dxg = Act.empty(n, t, h, w, c, da.dtype)
            ops.tdiff_bwd(da, dxg, False)
            return dxg, None
        ndf = self.mod.ndf
        if "stem" in ctx:
            if need_dx or need_dw:
                return self.merged.backward(ctx["stem"], da, sink, need_dx, need_dw)
            return None, None
        dxg = dxc = None
        if need_dx:
            xg, xc = ctx["stem_g"]["x"], ctx["stem_c"]["x"]
            dxg = Act.empty(xg.n, xg.t, xg.h, xg.w, xg.c, xg.dtype)
            dxc = Act.empty(xc.n, xc.t, xc.h, xc.w, xc.c, xc.dtype)
        if need_dx or need_dw:
            self.stem_g.backward(ctx["stem_g"], da.ch(ndf // 2, ndf), sink, dx_out=dxg, need_dw=need_dw)
            self.stem_c.backward(ctx["stem_c"], da.ch(0, ndf // 2), sink, dx_out=dxc, need_dw=need_dw)
        return dxg, dxc
