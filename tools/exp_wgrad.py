#!/usr/bin/env python
"""Timing experiments on single wgrad_tc launches (GPU box): partial-sum kernel alone (dcv_wgrad_partial), with TMA
loads switched off after the first ring fill (DCV_TC_DBG=3), and the full call including the split reduction."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcvgan_b200 import ops, require_device  # noqa: E402
from dcvgan_b200._lib import IMPL_TC, lib, check  # noqa: E402

B = 32
LAYERS = {
    # name: (kind, cin, cout, k, s, p, N, in_spatial)
    "vdis_main1": ("conv", 64, 128, (4, 4, 4), (1, 2, 2), (0, 1, 1), B, (13, 32, 32)),
    "vdis_main5": ("conv", 128, 256, (4, 4, 4), (1, 2, 2), (0, 1, 1), B, (10, 16, 16)),
    "vdis_stem": ("conv", 16, 64, (4, 4, 4), (1, 2, 2), (0, 1, 1), B, (16, 64, 64)),
    "down2": ("conv", 128, 256, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 16, 16)),
    "up3": ("convT", 512, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 8, 8)),
    "up2": ("convT", 512, 256, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 4, 4)),
    "up5": ("convT", 128, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 32, 32)),
    "down0": ("conv", 64, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 64, 64)),
    "down1": ("conv", 64, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 32, 32)),
    "outconv": ("convT", 128, 16, (1, 3, 3), (1, 1, 1), (0, 1, 1), 512, (1, 64, 64)),
    "inconv": ("conv", 16, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 512, (1, 64, 64)),
    "fold_stem": ("conv", 16, 64, (4, 4, 1), (1, 2, 1), (0, 1, 0), B, (16, 64, 32)),
    "fold_inconv": ("conv", 16, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 512, (1, 64, 64)),
    "fold_outconv": ("conv", 128, 32, (1, 1, 1), (1, 1, 1), (0, 0, 0), 512, (1, 64, 64)),
}


def timeit(fn, flush):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    require_device()
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
    print("| layer | GF | full ms | partial-kernel ms | partial, no loads ms | TFLOP/s (full) | ws MB |")
    for name in (sys.argv[1:] or list(LAYERS)):
        kind, cin, cout, k, s, p, n, sp = LAYERS[name]
        spec = ops.ConvSpec(kind, cin, cout, k, s, p)
        dt = torch.bfloat16
        x = ops.Act.empty(n, *sp, cin, dt)
        osp = spec.out_spatial(sp)
        y = ops.Act.empty(n, *osp, cout, dt)
        x.base.normal_()
        y.base.normal_()
        g = spec.geom(n, sp, x.cp, y.cp)
        taps = k[0] * k[1] * k[2]
        dw = torch.empty((cout, cin, taps) if kind == "conv" else (cin, cout, taps), device="cuda")
        xl, xs = (x, y) if kind == "conv" else (y, x)
        m_s = n * (osp if kind == "conv" else sp)[0] * (osp if kind == "conv" else sp)[1] * (osp if kind == "conv" else sp)[2]
        flops = 2.0 * m_s * taps * cin * cout
        nbytes = lib().dcv_wgrad_workspace_bytes(C.byref(g), IMPL_TC)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        full = lambda: ops.wgrad(spec, g, xl, xs, dw, False, IMPL_TC)
        part = lambda: check(lib().dcv_wgrad_partial(C.byref(g), IMPL_TC, 1, xl.ptr, xl.ld, xs.ptr, xs.ld, ws.data_ptr(), nbytes, st))
        os.environ.pop("DCV_TC_DBG", None)
        t_full, t_part = timeit(full, flush), timeit(part, flush)
        os.environ["DCV_TC_DBG"] = "3"
        t_nold = timeit(part, flush)
        os.environ.pop("DCV_TC_DBG", None)
        print(f"| {name} | {flops / 1e9:.1f} | {t_full:.3f} | {t_part:.3f} | {t_nold:.3f} | {flops / t_full / 1e9:.0f} | {nbytes / 2**20:.0f} |", flush=True)


if __name__ == "__main__":
    main()
