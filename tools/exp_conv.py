#!/usr/bin/env python
"""Timing experiments on single conv_tc launches (GPU box): which resource bounds the kernel?

For every probe layer the same launch is timed with the debug knobs of csrc/tc_conv.cu:
  base            - the production configuration
  noA / noB / noAB- DCV_TC_DBG=1/2/3: the producer stops issuing A / B / both TMA loads after the first ring fill
                    (wrong numbers, right MMA count) -> how much of the time is operand delivery
  mt1             - one M tile per work item;  nopersist - the multi-CTA (non-persistent) kernel
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcvgan_b200 import ops, require_device  # noqa: E402
from dcvgan_b200._lib import IMPL_TC  # noqa: E402
from tools.probe_layers import LAYERS  # noqa: E402

LAYERS.update({
    "down1_fwd": ("conv", 64, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 32, 32), "fwd"),
    "down2_fwd": ("conv", 128, 256, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 16, 16), "fwd"),
    "up4_fwd": ("convT", 256, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 16, 16), "fwd"),
    "up3_fwd": ("convT", 512, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 8, 8), "fwd"),
    "vdis_main5_fwd": ("conv", 128, 256, (4, 4, 4), (1, 2, 2), (0, 1, 1), 32, (10, 16, 16), "fwd"),
    "up5_dgrad": ("convT", 128, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 32, 32), "dgrad"),
    "outconv_dgrad": ("convT", 128, 3, (1, 3, 3), (1, 1, 1), (0, 1, 1), 512, (1, 64, 64), "dgrad"),
    "vdis_stem_fwd": ("conv", 4, 64, (4, 4, 4), (1, 2, 2), (0, 1, 1), 32, (16, 64, 64), "fwd"),
    "vdis_stem_dgrad": ("conv", 4, 64, (4, 4, 4), (1, 2, 2), (0, 1, 1), 32, (16, 64, 64), "dgrad"),
    "ggen_last_fwd": ("convT", 64, 1, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 32, 32), "fwd"),
    "up4_dgrad": ("convT", 256, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 16, 16), "dgrad"),
    "up3_dgrad": ("convT", 512, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 8, 8), "dgrad"),
    "down0_dgrad": ("conv", 64, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 64, 64), "dgrad"),
    "down1_dgrad": ("conv", 64, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 32, 32), "dgrad"),
    "vdis_main5_dgrad": ("conv", 128, 256, (4, 4, 4), (1, 2, 2), (0, 1, 1), 32, (10, 16, 16), "dgrad"),
    "ggen_main9_fwd": ("convT", 128, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 16, 16), "fwd"),
    "ggen_main6_fwd": ("convT", 256, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 512, (1, 8, 8), "fwd"),
    # tap-folded formulations of the image-like layers (taps moved into the tiny channel dimension by a cheap pre-pass)
    "fold_stem_fwd": ("conv", 16, 64, (4, 4, 1), (1, 2, 1), (0, 1, 0), 32, (16, 64, 32), "fwd"),
    "fold_stem_dgrad": ("conv", 16, 64, (4, 4, 1), (1, 2, 1), (0, 1, 0), 32, (16, 64, 32), "dgrad"),
    "fold_inconv_fwd": ("conv", 16, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 512, (1, 64, 64), "fwd"),
    "fold_inconv_dgrad": ("conv", 16, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 512, (1, 64, 64), "dgrad"),
    "fold_outconv_dgrad": ("conv", 32, 128, (1, 1, 1), (1, 1, 1), (0, 0, 0), 512, (1, 64, 64), "fwd"),
    "fold_outconv_fwd": ("conv", 128, 32, (1, 1, 1), (1, 1, 1), (0, 0, 0), 512, (1, 64, 64), "fwd"),
})
VARIANTS_FULL = [("base", {}), ("noA", {"DCV_TC_DBG": "1"}), ("noB", {"DCV_TC_DBG": "2"}), ("noAB", {"DCV_TC_DBG": "3"}),
            ("mt1", {"DCV_TC_MT1": "1"}), ("nopersist", {"DCV_TC_NOPERSIST": "1"}), ("nostore", {"DCV_TC_DBG": "4"}),
            ("noepi+noAB", {"DCV_TC_DBG": "11"}), ("nohalo", {"DCV_TC_NOHALO": "1"})]


VARIANTS = [("base", {}), ("mt1", {"DCV_TC_MT": "1"}), ("mt2", {"DCV_TC_MT": "2"}), ("mt4", {"DCV_TC_MT": "4"}),
            ("nohalo", {"DCV_TC_NOHALO": "1"}), ("nohalo,mt1", {"DCV_TC_NOHALO": "1", "DCV_TC_MT": "1"}),
            ("nohalo,mt2", {"DCV_TC_NOHALO": "1", "DCV_TC_MT": "2"}), ("nohalo,mt4", {"DCV_TC_NOHALO": "1", "DCV_TC_MT": "4"})]
if os.environ.get("EXP_FULL"):
    VARIANTS = VARIANTS_FULL


def build(name):
    kind, cin, cout, k, s, p, n, sp, op = LAYERS[name]
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    dt = torch.bfloat16
    x = ops.Act.empty(n, *sp, cin, dt)
    osp = spec.out_spatial(sp)
    y = ops.Act.empty(n, *osp, cout, dt)
    x.base.normal_()
    y.base.normal_()
    g = spec.geom(n, sp, x.cp, y.cp)
    w = (torch.randn((cout, cin, *k) if kind == "conv" else (cin, cout, *k[1:]), device="cuda") * 0.02).contiguous()
    d = spec.fwd_dir if op == "fwd" else spec.bwd_dir
    src, dst = (x, y) if op == "fwd" else (y, x)
    wp = ops.pack_weight(spec, g, d, IMPL_TC, w)
    taps = k[0] * k[1] * k[2]
    m_out = n * osp[0] * osp[1] * osp[2]
    flops = 2.0 * m_out * taps * cin * cout
    return (lambda: ops.conv(g, d, IMPL_TC, src.padded_to(src.cp), wp, dst.padded_to(dst.cp))), flops


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
    names = sys.argv[1:] or ["up4_dgrad", "up3_dgrad", "down0_dgrad", "down1_dgrad", "vdis_main5_dgrad", "ggen_main9_fwd", "ggen_main6_fwd","vdis_main1_fwd", "vdis_main1_dgrad", "vdis_main5_fwd", "up5_fwd", "up5_dgrad", "up4_fwd", "up3_fwd",
                             "down0_fwd", "down1_fwd", "down2_fwd", "outconv_fwd", "outconv_dgrad", "inconv_fwd", "vdis_stem_fwd", "vdis_stem_dgrad",
                             "ggen_last_fwd"]
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
    print("| layer | GF | " + " | ".join(v for v, _ in VARIANTS) + " |  (ms ; TFLOP/s of base)")
    for name in names:
        fn, flops = build(name)
        cells = []
        for tag, env in VARIANTS:
            for k in ("DCV_TC_DBG", "DCV_TC_CPS", "DCV_TC_MT1", "DCV_TC_NOPERSIST", "DCV_TC_NOHALO", "DCV_TC_MT"):
                os.environ.pop(k, None)
            os.environ.update(env)
            cells.append(timeit(fn, flush))
        for k in ("DCV_TC_DBG", "DCV_TC_CPS", "DCV_TC_MT1", "DCV_TC_NOPERSIST", "DCV_TC_NOHALO", "DCV_TC_MT"):
            os.environ.pop(k, None)
        print(f"| {name} | {flops / 1e9:.1f} | " + " | ".join(f"{c:.3f}" for c in cells) + f" | {flops / cells[0] / 1e9:.0f} TF/s", flush=True)


if __name__ == "__main__":
    main()
