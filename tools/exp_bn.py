#!/usr/bin/env python
"""Achieved HBM bandwidth of the BatchNorm / elementwise kernels at the shapes of the training step (GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcvgan_b200 import ops, require_device  # noqa: E402
from dcvgan_b200._lib import ACT_LEAKY  # noqa: E402

SHAPES = [(512, 64, 64, 64), (512, 32, 32, 64), (512, 16, 16, 128), (512, 8, 8, 256), (32 * 10, 16, 16, 128), (512, 4, 4, 256)]


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
    print("| shape (N,H,W,C) | MB/tensor | stats ms (GB/s) | bn_act ms (GB/s) | bwd ms (GB/s, 2 passes: 4 reads + 1 write) | copy_ ms (GB/s) |")
    for n, h, w, c in SHAPES:
        dt = torch.bfloat16
        z = ops.Act.empty(n, 1, h, w, c, dt)
        z.base.normal_()
        a, da, dz = z.like(), z.like(), z.like()
        da.base.normal_()
        gamma = torch.ones(c, device="cuda")
        beta = torch.zeros(c, device="cuda")
        dg, db = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        mb = z.base.numel() * 2 / 1e6
        mean, invstd = ops.bn_batch_stats(z, 1e-5, 0.1, None, None)
        t_s = timeit(lambda: ops.bn_batch_stats(z, 1e-5, 0.1, None, None), flush)
        t_a = timeit(lambda: ops.bn_act(z, mean, invstd, gamma, beta, None, ACT_LEAKY, 0.2, a), flush)
        t_b = timeit(lambda: ops.bn_act_bwd(da, a, z, mean, invstd, gamma, beta, None, ACT_LEAKY, 0.2, dz, dg, db), flush)
        t_c = timeit(lambda: a.base.copy_(z.base), flush)
        print(f"| {n},{h},{w},{c} | {mb:.0f} | {t_s:.3f} ({mb / t_s:.0f}) | {t_a:.3f} ({2 * mb / t_a:.0f}) | {t_b:.3f} ({5 * mb / t_b:.0f}) | "
              f"{t_c:.3f} ({2 * mb / t_c:.0f}) |", flush=True)


if __name__ == "__main__":
    main()
