"""The two util.py contracts the hot path depends on (util.py:16-28,186-195) plus the eval-mode
sampling loop (util.py:251-322)."""
from typing import Any, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.init as init

from .generator import current_device  # noqa: F401


def init_weights(layer: Any):
    """N(0,0.02) for exactly nn.Conv2d / nn.ConvTranspose2d, N(1,0.02)/0 for exactly nn.BatchNorm2d; Conv3d,
    BatchNorm3d and GRUCell keep torch's defaults (util.py:186-195)."""
    if type(layer) in [nn.Conv2d, nn.ConvTranspose2d]:
        init.normal_(layer.weight.data, 0, 0.02)
    elif type(layer) in [nn.BatchNorm2d]:
        init.normal_(layer.weight.data, 1.0, 0.02)
        init.constant_(layer.bias.data, 0.0)


def videos_to_numpy(tensor: torch.Tensor) -> np.ndarray:
    """float [-1,1] (B,C,T,H,W) -> uint8 (util.py:59-79)"""
    t = tensor.detach().clamp(-1, 1)
    return ((t + 1) / 2 * 255).to(torch.uint8).cpu().numpy()


def generate_samples(ggen, cgen, num: int, batchsize: int = 20, with_geo: bool = True) -> Tuple[Any, np.ndarray]:
    """Eval-mode, no-grad sampling (util.py:251-322).  Returns (raw geometry float (num,C,T,H,W) clipped to [-1,1]
    or None, colour uint8 (num,3,T,H,W)).

    Runs the two generator plans back to back on channels-last buffers (BatchNorm with running statistics, Dropout off)
    without the module-boundary layout conversions, and converts the colour video to uint8 ON THE DEVICE
    (dcv_export_u8: clip / (v+1)/2*255 / truncate, the arithmetic of util.py:74-79), so a quarter of the bytes cross
    PCIe and nothing is converted on the host."""
    from . import _lib, engine, ops
    _lib.require_device()
    ggen.eval()
    cgen.eval()
    dev = current_device()
    ggen.to(dev)
    cgen.to(dev)
    gplan, cplan = engine.GGenPlan(ggen), engine.CGenPlan(cgen)
    dtype = ops.torch_dtype(ggen.precision)
    T, C = ggen.video_length, ggen.channel
    r = engine.rng()
    xg_b, xc_b = [], []
    for _ in range(0, num, batchsize):
        B = batchsize
        xg, _ = gplan.forward(B, False, dtype, r, save=False)                                 # (B*T,1,64,64,C)
        z = r.normal((B, cgen.dim_z))                                                         # generator.py:355-359
        zs = z.unsqueeze(1).repeat(1, T, 1).view(B * T, -1)
        xc, _ = cplan.forward(xg, zs, False, r, save=False)                                   # (B*T,1,64,64,3)
        out = torch.empty((B, 3, T, 64, 64), dtype=torch.uint8, device=dev)
        ops.export_u8(xc.reshape_nt(B, T), out)
        xc_b.append(out.cpu().numpy())
        if with_geo:
            g = torch.empty((B, C, T, 64, 64), dtype=torch.float32, device=dev)
            ops.from_channels_last(xg.reshape_nt(B, T), g)
            xg_b.append(np.clip(g.cpu().numpy(), -1, 1))
    xg = np.concatenate(xg_b)[:num] if with_geo else None
    return xg, np.concatenate(xc_b)[:num]
