#!/usr/bin/env python
"""Launch the image-side mma.sync kernels in isolation (for `ncu --set full -k regex:img_conv3x3`).
usage: probe_img.py [inconv_fwd inconv_bwd outconv_fwd outconv_dgrad outconv_wgrad]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcvgan_b200 import ops, require_device  # noqa: E402
from dcvgan_b200._lib import ACT_LEAKY, ACT_NONE, ACT_TANH  # noqa: E402

N = 512


def run(name):
    dt = torch.bfloat16
    if name.startswith("inconv"):
        spec = ops.ConvSpec("conv", 1, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        small = ops.Act.empty(N, 1, 64, 64, 1, dt)
        big = ops.Act.empty(N, 1, 64, 64, 64, dt)
        w = torch.randn(64, 1, 3, 3, device="cuda") * 0.1
    else:
        spec = ops.ConvSpec("convT", 128, 3, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        small = ops.Act.empty(N, 1, 64, 64, 3, dt)
        big = ops.Act.empty(N, 1, 64, 64, 128, dt)
        w = torch.randn(128, 3, 3, 3, device="cuda") * 0.1
    small.base.normal_()
    big.base.normal_()
    g = spec.geom(N, (1, 64, 64), *((small.cp, big.cp) if spec.kind == "conv" else (big.cp, small.cp)))
    dw = torch.empty_like(w)
    big2 = big.like()
    big2.base.normal_()
    fns = {
        "inconv_fwd": lambda: ops.img_conv_fwd(spec, g, small, w, big, ACT_LEAKY, 0.01),
        "inconv_bwd": lambda: ops.img_conv_bwd(spec, g, big2, big, small, w, ACT_LEAKY, 0.01, dw, False, small.like()),
        "outconv_fwd": lambda: ops.img_conv_scatter(spec, g, big, w, small, ACT_TANH, 0.0),
        "outconv_dgrad": lambda: ops.img_conv_fwd(spec, g, small, w, big, ACT_NONE, 0.0),
        "outconv_wgrad": lambda: ops.img_conv_bwd(spec, g, big, None, small, w, ACT_NONE, 0.0, dw, False, None),
    }
    fn = fns[name]
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms", flush=True)


if __name__ == "__main__":
    require_device()
    for name in (sys.argv[1:] or ["inconv_fwd", "inconv_bwd", "outconv_fwd", "outconv_dgrad", "outconv_wgrad"]):
        run(name)
