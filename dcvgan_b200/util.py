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
    or None, colour uint8 (num,3,T,H,W))."""
    ggen.eval()
    cgen.eval()
    xg_b, xc_b = [], []
    for _ in range(0, num, batchsize):
        with torch.no_grad():
            xg = ggen.sample_videos(batchsize)
            xc = cgen.forward_videos(xg)
        if with_geo:
            xg_b.append(np.clip(xg.cpu().numpy(), -1, 1))
        xc_b.append(videos_to_numpy(xc))
    xg = np.concatenate(xg_b)[:num] if with_geo else None
    return xg, np.concatenate(xc_b)[:num]
