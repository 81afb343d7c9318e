#!/usr/bin/env python
"""Launch a few named layers in isolation (for `ncu --set full -k regex:conv_tc|wgrad_tc`).

usage: probe_layers.py [name ...]   names: vdis_main1_fwd, inconv_fwd, up5_fwd, outconv_fwd, vdis_main1_wgrad, vdis_convc_wgrad
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcvgan_b200 import ops, require_device  # noqa: E402
from dcvgan_b200._lib import IMPL_TC  # noqa: E402

B = 32
LAYERS = {
    # name: (kind, cin, cout, k, s, p, N, in_spatial, op)
    "vdis_main1_fwd": ("conv", 64, 128, (4, 4, 4), (1, 2, 2), (0, 1, 1), B, (13, 32, 32), "fwd"),
    "vdis_main1_dgrad": ("conv", 64, 128, (4, 4, 4), (1, 2, 2), (0, 1, 1), B, (13, 32, 32), "dgrad"),
    "vdis_main1_wgrad": ("conv", 64, 128, (4, 4, 4), (1, 2, 2), (0, 1, 1), B, (13, 32, 32), "wgrad"),
    "vdis_convc_wgrad": ("conv", 3, 32, (4, 4, 4), (1, 2, 2), (0, 1, 1), B, (16, 64, 64), "wgrad"),
    "inconv_fwd": ("conv", 1, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), B * 16, (1, 64, 64), "fwd"),
    "up5_fwd": ("convT", 128, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), B * 16, (1, 32, 32), "fwd"),
    "outconv_fwd": ("convT", 128, 3, (1, 3, 3), (1, 1, 1), (0, 1, 1), B * 16, (1, 64, 64), "fwd"),
    "down0_fwd": ("conv", 64, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), B * 16, (1, 64, 64), "fwd"),
    "down0_wgrad": ("conv", 64, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), B * 16, (1, 64, 64), "wgrad"),
    "up5_wgrad": ("convT", 128, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), B * 16, (1, 32, 32), "wgrad"),
}


def run(name, reps=1):
    kind, cin, cout, k, s, p, n, sp, op = LAYERS[name]
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    dt = torch.bfloat16
    x = ops.Act.empty(n, *sp, cin, dt)
    osp = spec.out_spatial(sp)
    y = ops.Act.empty(n, *osp, cout, dt)
    x.base.normal_()
    y.base.normal_()
    g = spec.geom(n, sp, x.cp, y.cp)
    w = torch.randn((cout, cin, *k) if kind == "conv" else (cin, cout, *k[1:]), device="cuda") * 0.02
    w = w.reshape(w.shape[0], w.shape[1], -1).contiguous() if kind == "conv" and k[0] == 1 else w.contiguous()
    if op in ("fwd", "dgrad"):
        d = spec.fwd_dir if op == "fwd" else spec.bwd_dir
        src, dst = (x, y) if op == "fwd" else (y, x)
        wp = ops.pack_weight(spec, g, d, IMPL_TC, w)
        fn = lambda: ops.conv(g, d, IMPL_TC, src.padded_to(src.cp), wp, dst.padded_to(dst.cp))
    else:
        dw = torch.empty_like(w)
        xl, xs = (x, y) if kind == "conv" else (y, x)
        fn = lambda: ops.wgrad(spec, g, xl.padded_to(xl.cp), xs.padded_to(xs.cp), dw, False, IMPL_TC)
    for _ in range(reps):
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
    for name in (sys.argv[1:] or list(LAYERS)):
        run(name)
