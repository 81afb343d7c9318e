"""Drop-in replacements for the reference's src/generator.py classes.

Same class names, constructor signatures, attributes, nn.Module parameter tree / state_dict keys
(generator.py:37-80,158-353) and public methods (sample_videos, make_hidden, forward,
forward_videos, __str__).  The submodules are kept as *parameter containers* only - so that
util.init_weights' exact-type test, optim.Adam(model.parameters()), state_dict exchange with the
reference and whole-module pickling keep working - while every forward/backward runs through the
sm_100a kernels of libdcvgan_b200.so (engine.GGenPlan / engine.CGenPlan).  No CPU path exists.
"""
import json
from typing import List

import torch
import torch.nn as nn

from . import _lib, engine, get_precision, ops
from .ops import Act


def current_device() -> torch.device:
    """util.py:16-28"""
    return torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")


def _needs_grad(tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _check_cuda(mod):
    _lib.require_device()
    p = next(mod.parameters())
    if not p.is_cuda:
        raise _lib.DcvError(f"{type(mod).__name__} parameters are on {p.device}; dcvgan_b200 only computes on a B200 "
                            "(call .to('cuda:0') as trainer.py:235-237 does) - there is no CPU fallback")


class _GGenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, batchsize, *params):
        plan = engine.GGenPlan(mod)
        save = any(ctx.needs_input_grad)      # grad mode is off inside Function.forward; this is the autograd-aware flag
        dtype = ops.torch_dtype(mod.precision)
        out, pctx = plan.forward(batchsize, mod.training, dtype, engine.rng(), save)
        B, T, C = batchsize, mod.video_length, mod.channel
        mem = torch.empty((B, T, C, 64, 64), dtype=torch.float32, device="cuda")
        video = mem.permute(0, 2, 1, 3, 4)       # (B,C,T,H,W) view of (B,T,C,H,W) memory, generator.py:136-139
        ops.from_channels_last(out.reshape_nt(B, T), video)
        ctx.plan, ctx.pctx, ctx.params, ctx.dtype = plan, pctx, params, dtype
        return video

    @staticmethod
    def backward(ctx, dvideo):
        mod = ctx.plan.mod
        B, T, C = dvideo.shape[0], mod.video_length, mod.channel
        dout = Act.empty(B, T, 64, 64, C, ctx.dtype)
        ops.to_channels_last(dvideo.float(), dout)
        sink = engine.GradSink()
        ctx.plan.backward(ctx.pctx, dout.reshape_nt(B * T, 1), sink)
        return (None, None) + tuple(sink.grads.get(p) for p in ctx.params)


class GeometricVideoGenerator(nn.Module):
    """The geometric information video generator (generator.py:11-155)."""

    def __init__(self, dim_z_content: int, dim_z_motion: int, channel: int, geometric_info: str, ngf: int = 64,
                 video_length: int = 16):
        super(GeometricVideoGenerator, self).__init__()
        self.dim_z_content = dim_z_content
        self.dim_z_motion = dim_z_motion
        self.channel = channel
        self.geometric_info = geometric_info
        self.video_length = video_length
        self.ngf = ngf
        dim_z = dim_z_motion + dim_z_content
        self.dim_z = dim_z

        self.recurrent: nn.Module = nn.GRUCell(dim_z_motion, dim_z_motion)
        chans = [dim_z, ngf * 8, ngf * 4, ngf * 2, ngf]
        modules: List[nn.Module] = []
        for i in range(4):
            modules += [nn.ConvTranspose2d(chans[i], chans[i + 1], 4, 1 if i == 0 else 2, 0 if i == 0 else 1, bias=False),
                        nn.BatchNorm2d(chans[i + 1]), nn.ReLU(inplace=True)]
        modules.append(nn.ConvTranspose2d(ngf, self.channel, 4, 2, 1, bias=False))
        modules.append(nn.Softmax(dim=1) if self.geometric_info == "segmentation" else nn.Tanh())
        self.main = nn.Sequential(*modules)

        self.device = current_device()
        self.precision = get_precision()

    def sample_videos(self, batchsize: int) -> torch.Tensor:
        """(batchsize, channel, video_length, 64, 64) in [-1,1] (softmax probabilities for segmentation),
        connected to the parameters through autograd; memory layout (B,T,C,H,W) like generator.py:118-141."""
        _check_cuda(self)
        return _GGenFn.apply(self, batchsize, *self.parameters())

    def __str__(self, name: str = "ggen") -> str:
        return json.dumps({name: {"dim_zc": self.dim_z_content, "dim_zm": self.dim_z_motion, "channel": self.channel,
                                  "geometric_info": self.geometric_info, "vlen": self.video_length, "ngf": self.ngf}})


class Inconv(nn.Module):
    """generator.py:158-182 (parameter container)"""

    def __init__(self, in_ch: int, out_ch: int):
        super(Inconv, self).__init__()
        self.main = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=1, padding=1, bias=False),
                                  nn.LeakyReLU(inplace=True))


class DownBlock(nn.Module):
    """generator.py:185-216 (parameter container)"""

    def __init__(self, in_ch: int, out_ch: int, dropout=False):
        super(DownBlock, self).__init__()
        layers = [nn.Conv2d(in_ch, out_ch, kernel_size=4, stride=2, padding=1, bias=False), nn.BatchNorm2d(out_ch),
                  nn.LeakyReLU(0.2, inplace=True)]
        if dropout:
            layers.insert(2, nn.Dropout2d(0.5, inplace=True))
        self.main = nn.Sequential(*layers)


class UpBlock(nn.Module):
    """generator.py:219-253 (parameter container)"""

    def __init__(self, in_ch: int, out_ch: int, dropout: bool = False):
        super(UpBlock, self).__init__()
        layers = [nn.ConvTranspose2d(in_ch, out_ch, kernel_size=4, stride=2, padding=1, bias=False),
                  nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True)]
        if dropout:
            layers.insert(2, nn.Dropout2d(0.5, inplace=True))
        self.main = nn.Sequential(*layers)


class Outconv(nn.Module):
    """generator.py:256-282 (parameter container)"""

    def __init__(self, in_ch: int, out_ch: int):
        super(Outconv, self).__init__()
        self.main = nn.Sequential(nn.ConvTranspose2d(in_ch, out_ch, kernel_size=3, stride=1, padding=1, bias=False),
                                  nn.Tanh())


class _CGenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, z, *params):
        plan = engine.CGenPlan(mod)
        need_dx = ctx.needs_input_grad[1]
        save = any(ctx.needs_input_grad)
        dtype = ops.torch_dtype(mod.precision)
        N, C, H, W = x.shape
        xa = Act.empty(N, 1, H, W, C, dtype)
        ops.to_channels_last(x.float(), xa)
        out, pctx = plan.forward(xa, z.reshape(N, -1).float(), mod.training, engine.rng(), save)
        y = torch.empty((N, 3, H, W), dtype=torch.float32, device="cuda")
        ops.from_channels_last(out, y)
        ctx.plan, ctx.pctx, ctx.params, ctx.dtype, ctx.need_dx = plan, pctx, params, dtype, need_dx
        return y

    @staticmethod
    def backward(ctx, dy):
        N, _, H, W = dy.shape
        dout = Act.empty(N, 1, H, W, 3, ctx.dtype)
        ops.to_channels_last(dy.float(), dout)
        sink = engine.GradSink()
        dxa = ctx.plan.backward(ctx.pctx, dout, sink, need_dx=ctx.need_dx)
        dx = None
        if ctx.need_dx and dxa is not None:
            dx = torch.empty((N, dxa.c, H, W), dtype=torch.float32, device="cuda")
            ops.from_channels_last(dxa, dx)
        return (None, dx, None) + tuple(sink.grads.get(p) for p in ctx.params)


class ColorVideoGenerator(nn.Module):
    """The color video generator (generator.py:285-448)."""

    def __init__(self, in_ch: int, dim_z: int, geometric_info: str, ngf: int = 64, video_length: int = 16):
        super(ColorVideoGenerator, self).__init__()
        self.in_ch = in_ch
        self.out_ch = 3
        self.dim_z = dim_z
        self.geometric_info = geometric_info
        self.ngf = ngf

        self.inconv = Inconv(in_ch, ngf * 1)
        self.down_blocks = nn.ModuleList([DownBlock(ngf * 1, ngf * 1), DownBlock(ngf * 1, ngf * 2),
                                          DownBlock(ngf * 2, ngf * 4), DownBlock(ngf * 4, ngf * 4),
                                          DownBlock(ngf * 4, ngf * 4), DownBlock(ngf * 4, ngf * 4)])
        self.up_blocks = nn.ModuleList([UpBlock(ngf * 4 + dim_z, ngf * 4, dropout=True), UpBlock(ngf * 8, ngf * 4, dropout=True),
                                        UpBlock(ngf * 8, ngf * 4), UpBlock(ngf * 8, ngf * 2), UpBlock(ngf * 4, ngf * 1),
                                        UpBlock(ngf * 2, ngf * 1)])
        self.outconv = Outconv(ngf * 2, self.out_ch)
        self.n_down_blocks = len(self.down_blocks)
        self.n_up_blocks = len(self.up_blocks)
        self.device = current_device()
        self.channel = 3
        self.video_length = video_length
        self.precision = get_precision()

    def make_hidden(self, batchsize: int) -> torch.Tensor:
        """generator.py:355-359"""
        z = engine.rng().normal((batchsize, self.dim_z))
        return z.unsqueeze(-1).unsqueeze(-1)  # (B, dim_z, 1, 1)

    def forward(self, x, z):
        """geometric frames (N,C,64,64) + colour code (N,dim_z,1,1) -> colour frames (N,3,64,64); generator.py:361-402"""
        _check_cuda(self)
        if x.shape[-1] != 64 or x.shape[-2] != 64:
            raise ValueError("ColorVideoGenerator works on 64x64 frames (six stride-2 stages down to 1x1)")
        return _CGenFn.apply(self, x, z, *self.parameters())

    def forward_videos(self, xs: torch.Tensor) -> torch.Tensor:
        """generator.py:404-435: one colour code per clip repeated over time; T folded into the batch."""
        B, C, T, H, W = xs.shape
        zs = self.make_hidden(B)
        zs = zs.unsqueeze(1).repeat(1, T, 1, 1, 1).view(B * T, -1, 1, 1)
        xs = xs.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
        ys = self(xs, zs)
        ys = ys.view(B, T, 3, H, W)
        return ys.permute(0, 2, 1, 3, 4)

    def __str__(self, name: str = "cgen") -> str:
        return json.dumps({name: {"in_ch": self.in_ch, "out_ch": self.out_ch, "dim_z": self.dim_z,
                                  "n_down_blocks": self.n_down_blocks, "n_up_blocks": self.n_up_blocks}})
