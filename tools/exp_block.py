#!/usr/bin/env python
"""Time engine.Block forwards of the tiny-Cout transposed convolutions with and without the tap-unrolled path (GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcvgan_b200  # noqa: E402
from dcvgan_b200 import engine, ops  # noqa: E402
from dcvgan_b200._lib import ACT_TANH  # noqa: E402


def main():
    dcvgan_b200.require_device()
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
    cases = [("outconv 128->3 k3s1 @64x64", 128, 3, 3, 1, 1, 64), ("ggen main.12 64->1 k4s2 @32x32", 64, 1, 4, 2, 1, 32)]
    for name, cin, cout, k, s, p, hw in cases:
        conv = torch.nn.ConvTranspose2d(cin, cout, k, s, p, bias=False).cuda()
        blk = engine.Block(engine.convT2d_spec(cin, cout, k, s, p), conv, None, ACT_TANH)
        x = ops.Act.empty(512, 1, hw, hw, cin, torch.bfloat16)
        x.base.normal_()
        osz = (hw - 1) * s - 2 * p + k
        out = ops.Act.empty(512, 1, osz, osz, cout, torch.bfloat16)
        res = {}
        for mode in (True, False):
            engine.TAP_UNROLL = mode
            engine.WCACHE = {}
            fn = lambda: blk.forward(x, out, True, engine.rng(), save=False)
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
            res[mode] = (float(np.median(ts)), out.base.float().clone())
        engine.WCACHE = None
        engine.TAP_UNROLL = True
        err = float((res[True][1] - res[False][1]).norm() / res[False][1].norm())
        print(f"{name}: tap-unrolled {res[True][0]:.3f} ms, direct {res[False][0]:.3f} ms, relative difference {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
