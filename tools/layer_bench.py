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
    tr.use_cuda_graph = False
    for _ in range(2):
        tr.iteration += 1
        tr.train_step(xc, xg)
    rows = bench.layer_table(tr, cfg, B)
    total = sum(r["ms"] * r["n"] for r in rows)
    print(f"{cfg_name} B={B}: conv-family total {total:.2f} ms/step over {sum(r['n'] for r in rows)} launches")
    print("| ms/step | ms | n | op | impl | L (T,H,W,C) | S (T,H,W,C) | k | GF | TFLOP/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        g = Geom(*r["key"])
        kind, direction, impl, flops, ms, n = r["kind"], r["dir"], r["impl"], r["flops"], r["ms"], r["n"]
        op = kind if kind != "conv" else ("gather" if direction == 0 else "scatter")
        print(f"| {ms * n:.3f} | {ms:.3f} | {n} | {op} | {'tc' if impl == IMPL_TC else ('direct' if impl < 0 else 'simt')} | {g.Tl},{g.Hl},{g.Wl},{g.wCl or g.Cl} | "
              f"{g.Ts},{g.Hs},{g.Ws},{g.wCs or g.Cs} | {g.kt}x{g.kh}x{g.kw} | {flops / 1e9:.1f} | {flops / ms / 1e9:.0f} |")


if __name__ == "__main__":
    main()
