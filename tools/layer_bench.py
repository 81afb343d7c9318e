#!/usr/bin/env python
"""Per-layer timing of every convolution-family launch of one training step (GPU box).

Records the conv / wgrad calls of one fused step (bench config, batch 32), then replays every distinct
(geometry, direction) in isolation with CUDA events and an L2 flush between launches, and prints time, useful
TFLOP/s (2*M*N*K with the real, unpadded channel counts) and the share of the step.
"""
import collections
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dcvgan_b200 import ops  # noqa: E402
from dcvgan_b200._lib import Geom, IMPL_TC  # noqa: E402


def main():
    cfg_name = sys.argv[1] if len(sys.argv) > 1 else "mug-depth"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    cfg = bench.make_cfg(cfg_name, B)
    tr = bench.build_trainer(cfg, "bf16")
    C_ = cfg["geometric_info"]["channel"]
    xc = torch.rand(B, 3, 16, 64, 64, device="cuda") * 2 - 1
    xg = torch.rand(B, C_, 16, 64, 64, device="cuda") * 2 - 1
    for _ in range(2):
        tr.iteration += 1
        tr.train_step(xc, xg)
    ops.TRACE = []
    tr.iteration += 1
    tr.train_step(xc, xg)
    trace, ops.TRACE = ops.TRACE, None
    torch.cuda.synchronize()
    count = collections.Counter((t[0], t[1], t[2], t[3]) for t in trace)
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
    rows = []
    for (kind, key, direction, impl), n in count.items():
        g = Geom(*key)
        taps = g.kt * g.kh * g.kw
        M = g.N * g.Ts * g.Hs * g.Ws
        wl, ws = (g.wCl or g.Cl), (g.wCs or g.Cs)
        flops = 2.0 * M * taps * wl * ws
        dt = torch.bfloat16
        L = ops.Act.empty(g.N, g.Tl, g.Hl, g.Wl, g.Cl, dt)
        S = ops.Act.empty(g.N, g.Ts, g.Hs, g.Ws, g.Cs, dt)
        L.base.normal_()
        S.base.normal_()
        if kind == "conv":
            nbytes = ops.lib().dcv_packed_weight_bytes(C.byref(g), direction, impl)
            wp = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
            x, y = (L, S) if direction == 0 else (S, L)
            fn = lambda: ops.conv(g, direction, impl, x, wp, y)
        else:
            spec = ops.ConvSpec("conv", wl, ws, (g.kt, g.kh, g.kw), (g.st, g.sh, g.sw), (g.pt, g.ph, g.pw))
            dw = torch.empty((ws, wl, taps), device="cuda")
            fn = lambda: ops.wgrad(spec, g, L, S, dw, False, impl)
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
        ms = float(np.median(ts))
        rows.append((ms * n, ms, n, kind, direction, impl, flops, key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    print(f"{cfg_name} B={B}: conv-family total {total:.2f} ms/step over {sum(r[2] for r in rows)} launches")
    print("| ms/step | ms | n | op | impl | L (T,H,W,C) | S (T,H,W,C) | k | GF | TFLOP/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for tot, ms, n, kind, direction, impl, flops, key in rows:
        g = Geom(*key)
        op = kind if kind == "wgrad" else ("gather" if direction == 0 else "scatter")
        print(f"| {tot:.3f} | {ms:.3f} | {n} | {op} | {'tc' if impl == IMPL_TC else 'simt'} | {g.Tl},{g.Hl},{g.Wl},{g.wCl or g.Cl} | "
              f"{g.Ts},{g.Hs},{g.Ws},{g.wCs or g.Cs} | {g.kt}x{g.kh}x{g.kw} | {flops / 1e9:.1f} | {flops / ms / 1e9:.0f} |")


if __name__ == "__main__":
    main()
