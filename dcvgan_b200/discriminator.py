"""Drop-in replacements for the reference's src/discriminator.py classes (Noise, Image/Video/Gradient
discriminators): same constructor signatures, attributes, state_dict keys and output shapes
(discriminator.py:11-346); compute goes through engine.DisPlan -> libdcvgan_b200.so.
"""
import json

import torch
import torch.nn as nn

from . import engine, get_precision, ops
from .generator import _check_cuda, _needs_grad, current_device
from .ops import Act


class Noise(nn.Module):
    """x + sigma * N(0,1), active in train and eval alike (discriminator.py:11-39).  Kept as a marker in the
    module tree; the addition itself is fused into the consumer block's input preparation."""

    def __init__(self, use_noise: bool, sigma: float = 0.2):
        super(Noise, self).__init__()
        self.use_noise = use_noise
        self.sigma = sigma
        self.device = current_device()

    def forward(self, x):
        if not self.use_noise:
            return x
        return _NoiseFn.apply(x, float(self.sigma))


class _NoiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sigma):
        _check_cuda_tensor(x)
        xa = Act.from_dense(x.detach().float().contiguous().view(1, 1, 1, -1, 1))
        out = xa.like()
        ops.add_noise(xa, engine.rng().normal(tuple(x.shape)).view(-1), sigma, out)
        return out.base.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        return dy, None


def _check_cuda_tensor(x):
    from . import _lib
    _lib.require_device()
    if not x.is_cuda:
        raise _lib.DcvError("dcvgan_b200 only computes on a B200; got a CPU tensor (no CPU fallback)")


class _DisFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, kind, xg, xc, *params):
        plan = engine.DisPlan(mod, kind)
        need_dx = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        save = any(ctx.needs_input_grad)
        dtype = ops.torch_dtype(mod.precision)

        def to_act(x):
            x5 = x if x.dim() == 5 else x.unsqueeze(2)
            n, c, t, h, w = x5.shape
            a = Act.empty(n, t, h, w, c, dtype)
            ops.to_channels_last(x5.float(), a)
            return a

        xga = to_act(xg)
        xca = to_act(xc) if kind != "gdis" else None
        logits, pctx = plan.forward(xga, xca, mod.training, engine.rng(), save)
        y = torch.empty((logits.n, 1, logits.t, logits.h, logits.w), dtype=torch.float32, device="cuda")
        ops.from_channels_last(logits, y)
        ctx.plan, ctx.pctx, ctx.params, ctx.dtype, ctx.need_dx = plan, pctx, params, dtype, need_dx
        ctx.shapes = (xg.shape, xc.shape)
        return y if xg.dim() == 5 else y[:, :, 0]

    @staticmethod
    def backward(ctx, dy):
        dy5 = dy if dy.dim() == 5 else dy.unsqueeze(2)
        n, _, t, h, w = dy5.shape
        dl = Act.empty(n, t, h, w, 1, ctx.dtype)
        ops.to_channels_last(dy5.float(), dl)
        sink = engine.GradSink()
        dxga, dxca = ctx.plan.backward(ctx.pctx, dl, sink, need_dx=ctx.need_dx, need_dw=True)
        grads = []
        for a, shape in ((dxga, ctx.shapes[0]), (dxca, ctx.shapes[1])):
            if a is None:
                grads.append(None)
                continue
            g = torch.empty(shape, dtype=torch.float32, device="cuda")
            ops.from_channels_last(a, g)
            grads.append(g)
        return (None, None, grads[0], grads[1]) + tuple(sink.grads.get(p) for p in ctx.params)


class _DisBase(nn.Module):
    KIND = ""

    def _init_common(self, ch1, ch2, use_noise, noise_sigma, ndf):
        self.ch1, self.ch2 = ch1, ch2
        self.use_noise = use_noise
        self.noise_sigma = noise_sigma
        self.ndf = ndf
        self.precision = get_precision()

    def forward(self, xg, xc):
        """(xg, xc) -> logits with every size-1 dimension squeezed away, like `self.main(h).squeeze()`."""
        _check_cuda(self)
        return _DisFn.apply(self, self.KIND, xg, xc, *self.parameters()).squeeze()

    def _describe(self, name):
        return json.dumps({name: {"ch_g": self.ch1, "ch_c": self.ch2, "ndf": self.ndf, "use_noise": self.use_noise,
                                  "noise_sigma": self.noise_sigma}})


class ImageDiscriminator(_DisBase):
    """The image discriminator (discriminator.py:42-140): logits (B, 4, 4) for 64x64 frames."""
    KIND = "idis"

    def __init__(self, ch1: int, ch2: int, use_noise: bool = False, noise_sigma: float = 0, ndf: int = 64):
        super(ImageDiscriminator, self).__init__()
        self._init_common(ch1, ch2, use_noise, noise_sigma, ndf)
        nz = lambda: Noise(use_noise, sigma=noise_sigma)
        lr = lambda: nn.LeakyReLU(0.2, inplace=True)
        self.conv_g = nn.Sequential(nz(), nn.Conv2d(ch1, ndf // 2, 4, 2, 1, bias=False), lr())
        self.conv_c = nn.Sequential(nz(), nn.Conv2d(ch2, ndf // 2, 4, 2, 1, bias=False), lr())
        self.main = nn.Sequential(nz(), nn.Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 2), lr(),
                                  nz(), nn.Conv2d(ndf * 2, ndf * 4, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 4), lr(),
                                  nz(), nn.Conv2d(ndf * 4, 1, 4, 2, 1, bias=False))
        self.device = current_device()

    def __str__(self, name: str = "idis") -> str:
        return self._describe(name)


def _c3(ci, co):
    return nn.Conv3d(ci, co, 4, stride=(1, 2, 2), padding=(0, 1, 1), bias=False)


class VideoDiscriminator(_DisBase):
    """The video discriminator (discriminator.py:143-244): logits (B, 4, 4, 4) for 16x64x64 clips."""
    KIND = "vdis"

    def __init__(self, ch1: int, ch2: int, use_noise: bool = False, noise_sigma: float = 0, ndf: int = 64):
        super(VideoDiscriminator, self).__init__()
        self._init_common(ch1, ch2, use_noise, noise_sigma, ndf)
        nz = lambda: Noise(use_noise, sigma=noise_sigma)
        lr = lambda: nn.LeakyReLU(0.2, inplace=True)
        self.conv_g = nn.Sequential(_c3(ch1, ndf // 2), lr())
        self.conv_c = nn.Sequential(_c3(ch2, ndf // 2), lr())
        self.main = nn.Sequential(nz(), _c3(ndf, ndf * 2), nn.BatchNorm3d(ndf * 2), lr(),
                                  nz(), _c3(ndf * 2, ndf * 4), nn.BatchNorm3d(ndf * 4), lr(),
                                  nz(), _c3(ndf * 4, 1))
        self.device = current_device()

    def __str__(self, name: str = "vdis") -> str:
        return self._describe(name)


class GradientDiscriminator(_DisBase):
    """The gradient discriminator (discriminator.py:247-346): temporal finite difference of xg only; xc unused.
    logits (B, 3, 4, 4)."""
    KIND = "gdis"

    def __init__(self, ch1: int, ch2: int, use_noise: bool = False, noise_sigma: float = 0, ndf: int = 64):
        super(GradientDiscriminator, self).__init__()
        self._init_common(ch1, ch2, use_noise, noise_sigma, ndf)
        nz = lambda: Noise(use_noise, sigma=noise_sigma)
        lr = lambda: nn.LeakyReLU(0.2, inplace=True)
        self.main = nn.Sequential(nz(), _c3(ch1, ndf), nn.BatchNorm3d(ndf), lr(),
                                  nz(), _c3(ndf, ndf * 2), nn.BatchNorm3d(ndf * 2), lr(),
                                  nz(), _c3(ndf * 2, ndf * 4), nn.BatchNorm3d(ndf * 4), lr(),
                                  nz(), _c3(ndf * 4, 1))
        self.device = current_device()

    def __str__(self, name: str = "vdis") -> str:  # the reference prints the key "vdis" here too (:335-338)
        return self._describe(name)
