#!/usr/bin/env python
"""Benchmark of the DCVGAN training step (BASELINE.json metric: train iters/sec, 16x64x64 clips, batch 32/GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config mug-depth]

One "step" = one full iteration of the reference's trainer.py:271-363 body (D-phase + G-phase, both updating)
on one batch of 32 synthetic clips per GPU.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# one visible device per rank so that the reference's hard-coded cuda:0 (util.py:25-26) stays valid
if "LOCAL_RANK" in os.environ and "DCV_KEEP_VISIBLE" not in os.environ:
    os.environ["CUDA_VISIBLE_DEVICES"] = os.environ["LOCAL_RANK"]
# the reference arm is the reference's CPU implementation: hide the GPUs so that its util.current_device() picks the CPU
if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"] or "--impl=reference" in sys.argv:
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

O = {"lr": 0.0002, "decay": 0.00001}
CONFIGS = {
    # config/mug-depth.yml normalised: C=1, ngf 64/64, ndf 64/64, adversarial loss, no noise, no gdis
    "mug-depth": {"geometric_info": {"name": "depth", "channel": 1}, "loss": "adversarial-loss", "seed": 0,
                  "ggen": {"dim_z_content": 40, "dim_z_motion": 10, "ngf": 64, "optimizer": O},
                  "cgen": {"dim_z_color": 10, "ngf": 64, "optimizer": O},
                  "idis": {"use_noise": False, "noise_sigma": 0.1, "ndf": 64, "optimizer": O},
                  "vdis": {"use_noise": False, "noise_sigma": 0.1, "ndf": 64, "optimizer": O},
                  "gdis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 32, "optimizer": O, "enabled": False}},
    # config/isogd-flow.yml: C=2, hinge loss, Noise 0.2 on idis/vdis, gdis ndf 32 without noise
    "isogd-flow": {"geometric_info": {"name": "optical-flow", "channel": 2}, "loss": "hinge-loss", "seed": 15,
                   "ggen": {"dim_z_content": 40, "dim_z_motion": 10, "ngf": 64, "optimizer": O},
                   "cgen": {"dim_z_color": 10, "ngf": 64, "optimizer": O},
                   "idis": {"use_noise": True, "noise_sigma": 0.2, "ndf": 64, "optimizer": O},
                   "vdis": {"use_noise": True, "noise_sigma": 0.2, "ndf": 64, "optimizer": O},
                   "gdis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 32, "optimizer": O, "enabled": True}},
    # config/surreal-depth1.yml: C=1, ggen.ngf 96, hinge loss, no noise, no gdis (num_gen_update 2 in the YAML; the bench
    # updates D and G every iteration)
    "surreal-depth1": {"geometric_info": {"name": "depth", "channel": 1}, "loss": "hinge-loss", "seed": 15,
                       "ggen": {"dim_z_content": 40, "dim_z_motion": 10, "ngf": 96, "optimizer": O},
                       "cgen": {"dim_z_color": 10, "ngf": 64, "optimizer": O},
                       "idis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 64, "optimizer": O},
                       "vdis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 64, "optimizer": O},
                       "gdis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 32, "optimizer": O, "enabled": False}},
    # config/surreal-segm.yml: 25-channel segmentation (softmax geometry, argmax remap), ggen.ngf 96, vdis.ndf 48, Noise 0.2
    "surreal-segm": {"geometric_info": {"name": "segmentation", "channel": 25}, "loss": "adversarial-loss", "seed": 15,
                     "ggen": {"dim_z_content": 40, "dim_z_motion": 10, "ngf": 96, "optimizer": O},
                     "cgen": {"dim_z_color": 10, "ngf": 64, "optimizer": O},
                     "idis": {"use_noise": True, "noise_sigma": 0.2, "ndf": 64, "optimizer": O},
                     "vdis": {"use_noise": True, "noise_sigma": 0.2, "ndf": 48, "optimizer": O},
                     "gdis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 32, "optimizer": O, "enabled": False}},
}


def make_cfg(name, batch):
    cfg = json.loads(json.dumps(CONFIGS[name]))
    cfg.update({"batchsize": batch, "video_length": 16, "image_size": 64, "num_gen_update": 1, "num_dis_update": 1,
                "n_epochs": 1, "log_interval": 10 ** 9, "snapshot_interval": 10 ** 9, "log_samples_interval": 10 ** 9,
                "evaluation_interval": 10 ** 9, "evaluation": {"batchsize": 50, "num_samples": 0, "metrics": []},
                "config_path": ""})
    return cfg


def step_flops(cfg, B):
    """Useful FLOPs of one iteration = 4 F_G + 8 F_D with F = sum of forward 2*M*N*K."""
    C, T = cfg["geometric_info"]["channel"], 16
    n = B * T
    g, c = cfg["ggen"]["ngf"], cfg["cgen"]["ngf"]

    def conv(m, nn_, k):
        return 2.0 * m * nn_ * k
    fg = conv(n, 16 * 8 * g, 50)
    for ci, co, hin in ((8 * g, 4 * g, 4), (4 * g, 2 * g, 8), (2 * g, g, 16), (g, C, 32)):
        fg += conv(n * hin * hin * 4, co, 4 * ci)
    fg += conv(n * 64 * 64, c, 9 * C)
    for ci, co, hout in ((c, c, 32), (c, 2 * c, 16), (2 * c, 4 * c, 8), (4 * c, 4 * c, 4), (4 * c, 4 * c, 2), (4 * c, 4 * c, 1)):
        fg += conv(n * hout * hout, co, 16 * ci)
    for ci, co, hin in ((4 * c + 10, 4 * c, 1), (8 * c, 4 * c, 2), (8 * c, 4 * c, 4), (8 * c, 2 * c, 8), (4 * c, c, 16), (2 * c, c, 32)):
        fg += conv(n * hin * hin * 4, co, 4 * ci)
    fg += conv(n * 64 * 64, 3, 9 * 2 * c)
    fd = 0.0
    d = cfg["idis"]["ndf"]
    fd += conv(B * 32 * 32, d // 2, 16 * C) + conv(B * 32 * 32, d // 2, 48) + conv(B * 256, 2 * d, 16 * d) + conv(B * 64, 4 * d, 32 * d) + conv(B * 16, 1, 64 * d)
    d = cfg["vdis"]["ndf"]
    fd += conv(B * 13 * 1024, d // 2, 64 * C) + conv(B * 13 * 1024, d // 2, 192) + conv(B * 10 * 256, 2 * d, 64 * d) + conv(B * 7 * 64, 4 * d, 128 * d) + conv(B * 4 * 16, 1, 256 * d)
    if cfg["gdis"].get("enabled", True):
        d = cfg["gdis"]["ndf"]
        fd += conv(B * 12 * 1024, d, 64 * C) + conv(B * 9 * 256, 2 * d, 64 * d) + conv(B * 6 * 64, 4 * d, 128 * d) + conv(B * 3 * 16, 1, 256 * d)
    return 4 * fg + 8 * fd, fg, fd


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.stop_flag = threading.Event()
        self.rows = []
        self.index = index

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None),
                "reasons": sorted(reasons), "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d["bf16_tflops_sustained"], d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def synthetic_host_batches(cfg, B, n, seed):
    """n batches of synthetic clips shaped like the dataset's (dataset.py:111-186): color U(-1,1), geometry per kind"""
    from oracle import dcvgan_oracle as orc
    gname = cfg["geometric_info"]["name"]
    out = []
    for i in range(n):
        xc, xg = orc.synthetic_batch(cfg, B, seed + i)
        out.append({"color": xc, gname: xg})
    return out


def reference_cpu_run(cfg_name, B, warmup, steps):
    """The reference's own CPU implementation of the step, all host threads, at the benchmarked batch size.

    kind "reference": the UNMODIFIED modules and Trainer.train() loop of raahii/dcvgan staged under baseline/_ref/src
    (tools/ref_harness.py; staged by __graft_entry__.build()).  kind "port": the pinned oracle restatement
    (oracle/dcvgan_oracle.py) when the staged sources are absent.  Returns (iters/s, s/iter, kind, threads, note)."""
    from tools import ref_harness as rh
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    cfg = make_cfg(cfg_name, B)
    if rh.reference_src() is not None:
        ref = rh.import_reference()
        batches = synthetic_host_batches(cfg, B, 1, 1000) * (warmup + steps)
        _, logger = rh.run_reference_trainer(ref, cfg, batches, device="cpu", seed=cfg["seed"])
        ips, spi = rh.timed_iters_per_s(logger, warmup, steps)
        return ips, spi, "reference", torch.get_num_threads(), "unmodified reference Trainer.train() (baseline/_ref/src)"
    from oracle import dcvgan_oracle as orc
    tr = orc.OracleTrainer(cfg, orc.init_all(cfg, cfg["seed"]))
    torch.manual_seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    xc, xg = orc.synthetic_batch(cfg, B, 1000)
    for _ in range(warmup):
        tr.step(xc, xg)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(xc, xg)
    spi = (time.perf_counter() - t0) / steps
    return 1.0 / spi, spi, "port", torch.get_num_threads(), "oracle port (reference sources not staged)"


def run_reference(args, cfg_name):
    """`--impl reference`: rank 0 times the reference's CPU training step at batch 32 (1 warm-up + <= 3 timed iterations)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 32
    ips, spi, kind, threads, note = reference_cpu_run(cfg_name, B, args.warmup, args.steps)
    sample = f"{args.steps} timed iteration(s) of the full step at batch {B} after {args.warmup} warm-up, fp32, {threads} threads; {note}"
    line = {"impl": "reference", "metric": "train iters/sec (16x64x64 clips, batch 32)", "value": ips, "unit": "iters/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spi * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config/{cfg_name}.yml (normalised) training step, batch 32 per GPU, 16x64x64, D and G both update every iteration"},
            "cpu_baseline": {"value": ips, "unit": "iters/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- CUDA arm
def build_trainer(cfg, precision):
    import dcvgan_b200
    from dcvgan_b200 import discriminator, engine, generator, loss, trainer, util
    dcvgan_b200.require_device()
    dcvgan_b200.set_precision(precision)
    engine.set_rng_mode("device")
    C, gname = cfg["geometric_info"]["channel"], cfg["geometric_info"]["name"]
    torch.manual_seed(cfg["seed"])
    models = {"ggen": generator.GeometricVideoGenerator(40, 10, C, gname, cfg["ggen"]["ngf"], 16),
              "cgen": generator.ColorVideoGenerator(C, 10, gname, cfg["cgen"]["ngf"], 16),
              "idis": discriminator.ImageDiscriminator(C, 3, cfg["idis"]["use_noise"], cfg["idis"]["noise_sigma"], cfg["idis"]["ndf"]),
              "vdis": discriminator.VideoDiscriminator(C, 3, cfg["vdis"]["use_noise"], cfg["vdis"]["noise_sigma"], cfg["vdis"]["ndf"])}
    if cfg["gdis"].get("enabled", True):
        models["gdis"] = discriminator.GradientDiscriminator(C, 3, cfg["gdis"]["use_noise"], cfg["gdis"]["noise_sigma"], cfg["gdis"]["ndf"])
    for m in models.values():
        m.apply(util.init_weights)
        m.cuda()
    opts = {k: torch.optim.Adam(m.parameters(), lr=cfg[k]["optimizer"]["lr"], betas=(0.5, 0.999),
                                weight_decay=cfg[k]["optimizer"]["decay"]) for k, m in models.items()}
    L = loss.AdversarialLoss() if cfg["loss"] == "adversarial-loss" else loss.HingeLoss()

    class _Log:
        path = "/tmp/dcv_bench_rank%s" % os.environ.get("RANK", "0")

        def __getattr__(self, n):
            return lambda *a, **k: None
    os.makedirs(_Log.path, exist_ok=True)
    trainer.Trainer.save_classobj = lambda self: None
    return trainer.Trainer(None, _Log(), models, opts, L, cfg)


def layer_table(tr, cfg, B, reps=5):
    """Every convolution-family launch of one training step, timed alone: records the conv / wgrad calls of one eager
    step, then replays each distinct (kernel, geometry, direction) with CUDA events and an L2 flush between launches.
    Returns rows {ms, n, kind, dir, impl, flops, key}; flops = 2*M*N*K with the real (unpadded) channel counts."""
    import collections
    import ctypes as C
    from dcvgan_b200 import ops
    from dcvgan_b200._lib import Geom
    Cg = cfg["geometric_info"]["channel"]
    xc = torch.rand(B, 3, 16, 64, 64, device="cuda") * 2 - 1
    xg = torch.rand(B, Cg, 16, 64, 64, device="cuda") * 2 - 1
    ops.TRACE = []
    tr.iteration += 1
    tr._eager_step(xc, xg, 3)
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
        elif kind == "wgrad":
            spec = ops.ConvSpec("conv", wl, ws, (g.kt, g.kh, g.kw), (g.st, g.sh, g.sw), (g.pt, g.ph, g.pw))
            dw = torch.empty((ws, wl, taps), device="cuda")
            fn = lambda: ops.wgrad(spec, g, L, S, dw, False, impl)
        elif kind in ("img_conv_fwd", "img_conv_bwd", "img_conv_wgrad", "img_conv_scatter"):
            spec = ops.ConvSpec("conv", wl, ws, (g.kt, g.kh, g.kw), (g.st, g.sh, g.sw), (g.pt, g.ph, g.pw))
            w = torch.randn((ws, wl, g.kh, g.kw), device="cuda") * 0.1
            xin = ops.Act.empty(g.N, 1, g.Hl, g.Wl, wl, dt)
            if kind == "img_conv_fwd":
                fn = lambda: ops.img_conv_fwd(spec, g, xin, w, S, 1, 0.01)
            elif kind == "img_conv_scatter":
                yo = ops.Act.empty(g.N, 1, g.Hl, g.Wl, wl, dt)
                fn = lambda: ops.img_conv_scatter(spec, g, S, w, yo, 2, 0.0)
            elif kind == "img_conv_wgrad":
                dw = torch.empty_like(w)
                fn = lambda: ops.img_conv_bwd(spec, g, S, None, xin, w, 0, 0.0, dw, False, None)
            else:
                dw = torch.empty_like(w)
                dxo = ops.Act.empty(g.N, 1, g.Hl, g.Wl, wl, dt)
                S2 = S.like()
                S2.base.normal_()
                fn = lambda: ops.img_conv_bwd(spec, g, S2, S, xin, w, 1, 0.01, dw, False, dxo)
                flops *= 2
        else:
            continue
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        rows.append({"ms": float(np.median(ts)), "n": n, "kind": kind, "dir": direction, "impl": impl, "flops": flops, "key": key})
    rows.sort(key=lambda r: -r["ms"] * r["n"])
    return rows


def family_of(row):
    from dcvgan_b200._lib import IMPL_TC
    if row["kind"] == "conv":
        return "conv_tc_pers_kernel (tcgen05 implicit-GEMM convolution: forward + data gradient, all template variants)" if row["impl"] == IMPL_TC else "conv_simt_kernel"
    if row["kind"] == "wgrad":
        return "wgrad_tc_kernel (tcgen05 weight gradient)" if row["impl"] == IMPL_TC else "wgrad_simt_kernel"
    return "img_conv3x3 (direct CUDA-core kernels, HBM-bound)"


def describe_layer(row):
    from dcvgan_b200._lib import Geom
    g = Geom(*row["key"])
    op = {"conv": "gather" if row["dir"] == 0 else "scatter"}.get(row["kind"], row["kind"])
    return f"{op} L({g.Tl},{g.Hl},{g.Wl},{g.wCl or g.Cl}) S({g.Ts},{g.Hs},{g.Ws},{g.wCs or g.Cs}) k{g.kt}x{g.kh}x{g.kw} N={g.N}"


def hbm_probe(B, iters=10):
    """bn_act (normalise + affine + ReLU) on the 512x64x64x64 bf16 U-Net tensor (268 MB at B=32): 2 B read + 2 B written
    per element, timed alone with an L2 flush between launches."""
    from dcvgan_b200 import ops
    from dcvgan_b200._lib import ACT_LEAKY
    n = B * 16
    z = ops.Act.empty(n, 1, 64, 64, 64, torch.bfloat16)
    z.base.normal_()
    out = z.like()
    c = 64
    mean, invstd = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        ops.bn_act(z, mean, invstd, gamma, beta, None, ACT_LEAKY, 0.0, out)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.bn_act(z, mean, invstd, gamma, beta, None, ACT_LEAKY, 0.0, out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.mean(ts))
    nbytes = 2.0 * z.rows * c * 2
    return {"kernel": "bn_act_bf16_kernel (cgen up_blocks.5 BatchNorm + ReLU, 512x64x64x64)", "ms": ms, "bytes": nbytes, "gbs": nbytes / ms / 1e6}


def cpu_baseline(cfg_name):
    """The reference's CPU step timed on this box's host cores on a bounded sample of the same workload (rank 0, N=1):
    1 warm-up + 2 timed iterations at batch 32 (about 15-25 s)."""
    ips, spi, kind, threads, note = reference_cpu_run(cfg_name, 32, 1, 2)
    return {"value": ips, "unit": "iters/s", "cores": threads, "kind": kind,
            "sample": f"2 timed iterations of the full step at batch 32 after 1 warm-up, fp32; {note}"}


def gpu_eager_baseline(cfg_name, B=32, warmup=5, steps=50):
    """the UNMODIFIED reference run eagerly on this B200 ("existing Blackwell kernels": ATen /
    cuDNN) - torch defaults (fp32 with cuDNN TF32 convolutions), and a fair fast variant (bf16 autocast + channels_last).
    Device-synchronised host timestamps around exactly `steps` iterations after `warmup`."""
    from tools import ref_harness as rh
    if rh.reference_src() is None:
        return {"unavailable": "reference sources not staged under baseline/_ref/src"}
    ref = rh.import_reference()
    cfg = make_cfg(cfg_name, B)
    batches = synthetic_host_batches(cfg, B, 1, 1000)
    dev_batches = [{k: v.cuda() for k, v in batches[0].items()}] * (warmup + steps)
    out = {"unit": "iters/s", "batch": B, "steps": steps, "warmup": warmup, "inputs": "resident on the device",
           "what": "unmodified reference modules + Trainer.train() on cuda:0 (PyTorch %s eager)" % torch.__version__}
    for name, kw in (("tf32_default", {}), ("bf16_autocast_channels_last", {"autocast_dtype": torch.bfloat16, "channels_last": True})):
        try:
            torch.cuda.empty_cache()
            _, logger = rh.run_reference_trainer(ref, cfg, dev_batches, device="cuda:0", seed=cfg["seed"], **kw)
            ips, spi = rh.timed_iters_per_s(logger, warmup, steps)
            out[name] = {"value": ips, "ms_per_step": spi * 1e3}
        except Exception as e:  # noqa: BLE001 - a failing variant must not take the benchmark line down
            out[name] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    torch.cuda.empty_cache()
    return out


def run_cuda(args, cfg_name):
    import torch.distributed as dist
    from dcvgan_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", device_id=torch.device("cuda:0"))
    B = 32
    cfg = make_cfg(cfg_name, B)
    for kv in args.tune:
        k, v = kv.split("=")
        _lib.check(_lib.lib().dcv_set_tuning(k.encode(), int(v)))
    tr = build_trainer(cfg, args.precision)
    torch.manual_seed(cfg["seed"] + rank)
    np.random.seed(cfg["seed"] + rank)
    C = cfg["geometric_info"]["channel"]
    gen = torch.Generator().manual_seed(cfg["seed"] + 1000 * rank)
    nbuf = 4
    host = [(torch.rand((B, 3, 16, 64, 64), generator=gen).mul_(2).sub_(1).pin_memory(),
             torch.rand((B, C, 16, 64, 64), generator=gen).mul_(2).sub_(1).pin_memory()) for _ in range(nbuf)]
    dev = [(a.cuda(), b.cuda()) for a, b in host]
    h2d = sum(t.numel() * 4 for t in host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = []

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        host_ms.append((time.perf_counter() - t0) * 1e3 / steps)   # time the host needs to issue one step
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    def resident_step(i):
        tr.iteration += 1
        xc, xg = dev[i % nbuf]
        tr.train_step(xc, xg)

    loss_host = torch.empty(4).pin_memory()

    def e2e_step(i):
        tr.iteration += 1
        xc, xg = host[i % nbuf]
        l = tr.train_step(xc, xg)      # pinned host batch: its H2D copy is inside the timed region (staged by the previous step's prefetch)
        nxc, nxg = host[(i + 1) % nbuf]
        tr.prefetch(nxc, nxg)          # next step's batch crosses PCIe on the copy stream while this step computes (Trainer.train does the same)
        loss_host.copy_(l, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the user reads the losses each step (trainer.py:326-328,363)

    for i in range(max(args.warmup, 3)):
        resident_step(i)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")) if "DCV_KEEP_VISIBLE" in os.environ else 0)
    if rank == 0:
        sampler.start()
    l0 = _lib.lib().dcv_launch_count() + tr.replayed_launches
    ms = timed(resident_step, args.steps)
    launches = _lib.lib().dcv_launch_count() + tr.replayed_launches - l0
    ms_e2e = timed(e2e_step, args.steps)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    flops, fg, fd = step_flops(cfg, B)
    burst, sustained, hbm, src = peaks()
    line = {"metric": "train iters/sec (16x64x64 clips, batch 32)", "value": world * 1e3 / ms, "unit": "iters/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"config/{cfg_name}.yml (normalised) training step, batch 32 per GPU, 16x64x64, D and G both update every iteration",
                       "l2": f"{nbuf} rotating input batches; per-step activation traffic >> 126 MB L2",
                       "parallelism": f"dp{world}", "cuda_graph": bool(tr.use_cuda_graph), "useful_tflop_per_step": flops / 1e12, "host_issue_ms_per_step": host_ms[0],
                       "step_tflops": flops / ms / 1e9, "step_frac_of_sustained_peak": flops / ms / 1e9 / sustained},
            "e2e": {"value": world * 1e3 / ms_e2e, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16},
            "gpu_launches": int(launches),
            "clocks": sampler.summary() if rank == 0 else None}
    rows = None
    if args.precision == "bf16" and not args.no_roofline:
        # one traced eager iteration on every rank (it contains the gradient all-reduce), replays on rank 0 only
        if rank == 0:
            rows = layer_table(tr, cfg, B)
        else:
            Cg = cfg["geometric_info"]["channel"]
            tr.iteration += 1
            tr._eager_step(torch.rand(B, 3, 16, 64, 64, device="cuda"), torch.rand(B, Cg, 16, 64, 64, device="cuda"), 3)
    if rank == 0:
        line["roofline_step"] = {"bound": "tensor", "achieved": flops / ms / 1e9, "peak": sustained, "unit": "TFLOP/s",
                                 "frac": flops / ms / 1e9 / sustained, "what": "useful FLOPs of the whole iteration (4 F_G + 8 F_D) / ms_per_step",
                                 "peak_source": f"{src} sustained"}
        if rows:
            fams = {}
            for r in rows:
                f = fams.setdefault(family_of(r), {"ms": 0.0, "flops": 0.0, "n": 0})
                f["ms"] += r["ms"] * r["n"]
                f["flops"] += r["flops"] * r["n"]
                f["n"] += r["n"]
            name, top = max(fams.items(), key=lambda kv: kv[1]["ms"])          # kernel with the largest share of the step
            big = next(r for r in rows if family_of(r) == name)                  # its most expensive layer
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
            if os.path.exists(tpath):   # dram__bytes_read+write per launch from the committed `ncu --set full` capture of that layer
                traffic = json.load(open(tpath)).get(describe_layer(big), {}).get("dram_bytes_per_launch")
            line["roofline"] = {"bound": "tensor", "achieved": top["flops"] / top["ms"] / 1e9, "peak": burst, "unit": "TFLOP/s",
                                "frac": top["flops"] / top["ms"] / 1e9 / burst, "traffic": traffic, "kernel": name,
                                "launches_per_step": top["n"], "ms_per_step": top["ms"], "share_of_step": top["ms"] / ms,
                                "algorithmic_flops_per_launch": top["flops"] / top["n"], "ms_per_launch": top["ms"] / top["n"],
                                "how": "every launch of the kernel in one iteration replayed alone (CUDA events, L2 flushed between launches); achieved = sum of algorithmic FLOPs / sum of launch times",
                                "largest_launch": {"layer": describe_layer(big), "ms": big["ms"], "n_per_step": big["n"], "tflops": big["flops"] / big["ms"] / 1e9,
                                                   "frac": big["flops"] / big["ms"] / 1e9 / burst},
                                "peak_source": f"{src} burst (kernels timed alone)"}
            line["kernel_families"] = {k: {"ms_per_step": round(v["ms"], 4), "launches": v["n"], "tflops": round(v["flops"] / v["ms"] / 1e9, 1),
                                           "share_of_step": round(v["ms"] / ms, 4)} for k, v in fams.items()}
            hb = hbm_probe(B)
            line["roofline_hbm"] = {"bound": "hbm", "achieved": hb["gbs"], "peak": hbm, "unit": "GB/s", "frac": hb["gbs"] / hbm, "traffic": None,
                                    "kernel": hb["kernel"], "algorithmic_bytes_per_launch": hb["bytes"], "ms_per_launch": hb["ms"]}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(cfg_name)
        if world == 1 and not args.no_eager:
            line["gpu_eager_baseline"] = gpu_eager_baseline(cfg_name)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured CUDA graphs hold NCCL work; tearing the communicator down under them can block (observed: the
        # 2-GPU run printed its line and then sat in destroy_process_group until the timeout).  Drop the graphs, agree
        # that everyone is done, and leave without the collective teardown.
        tr._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="mug-depth", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the gpu_eager_baseline leg (unmodified reference on cuda:0)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-layer replay behind the roofline objects")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE", help="dcv_set_tuning switch for A/B runs, e.g. no_pdl=1")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 3:
            args.steps = 3      # bounded sample: a batch-32 CPU iteration takes ~10-20 s
        args.warmup = min(args.warmup, 1)
        run_reference(args, args.config)
    else:
        run_cuda(args, args.config)


if __name__ == "__main__":
    main()
