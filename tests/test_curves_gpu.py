"""bf16 loss-curve gate of the north star: generator and discriminator loss curves over 200 steps at bf16 stay within
a stated tolerance of the fp32 curves (same seeds, same host-generated noise, same synthetic batches).

The fp32 path is itself gated against the reference (test_nets_gpu.py: golden runs, losses <= 1e-3, gradient cosine
>= 0.999), so it stands in for the reference over the 200 steps; a shorter prefix is also checked directly against the
CPU oracle.  GAN trajectories are chaotic point-wise: an fp32 run whose initial weights are perturbed by 1e-6
(relative) drifts from the unperturbed fp32 run by up to ~0.46 in the running mean of loss_vdis within 200 steps
(measured, profiles/r1_curves.md).  Stated tolerance, on running means over 20 iterations:
  * first 30 windows (before the trajectories decorrelate): |bf16 - fp32| <= max(5 % of the fp32 value, 0.05);
  * all 181 windows: |bf16 - fp32| <= max(2 x the largest drift among three perturbed fp32 control runs (initial
    weights scaled by 1+1e-6, 1+1e-3, 1-1e-3), 0.05) per curve.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dcvgan_oracle as orc  # noqa: E402
from test_nets_gpu import _Logger, _mods, build_models, small_cfg  # noqa: E402

STEPS, WINDOW = 200, 20


def _run(cfg, init, precision, steps, tmp_path):
    dcv, _, _, loss_mod, trainer_mod, _, engine = _mods()
    models = build_models(cfg, init, precision)
    engine.set_rng_mode("cpu_parity")
    opts = {k: torch.optim.Adam(m.parameters(), lr=cfg[k]["optimizer"]["lr"], betas=(0.5, 0.999),
                                weight_decay=cfg[k]["optimizer"]["decay"]) for k, m in models.items()}
    L = loss_mod.AdversarialLoss() if cfg["loss"] == "adversarial-loss" else loss_mod.HingeLoss()
    trainer_mod.Trainer.save_classobj = lambda self: None
    tr = trainer_mod.Trainer(None, _Logger(tmp_path), models, opts, L, dict(cfg, config_path=""))
    torch.manual_seed(123)
    np.random.seed(123)
    out = []
    for it in range(steps):
        xc, xg = orc.synthetic_batch(cfg, cfg["batchsize"], 5000 + it % 8)
        tr.iteration += 1
        out.append(tr.train_step(xc.cuda(), xg.cuda()))
    return torch.stack(out).cpu().numpy()


def _running_mean(a, w):
    c = np.cumsum(np.insert(a, 0, 0.0, axis=0), axis=0)
    return (c[w:] - c[:-w]) / w


def test_bf16_loss_curves_track_fp32(tmp_path):
    cfg = small_cfg("optical-flow", 2, "hinge-loss", noise=True, ngf=16, ndf=16)
    cfg["batchsize"] = 4
    init = orc.init_all(cfg, 9)
    ref = _run(cfg, init, "fp32", STEPS, tmp_path)
    got = _run(cfg, init, "bf16", STEPS, tmp_path)
    assert np.isfinite(ref).all() and np.isfinite(got).all()
    mr, mg = _running_mean(ref, WINDOW), _running_mean(got, WINDOW)
    dev = np.abs(mg - mr)
    names = ("loss_idis", "loss_vdis", "loss_gdis", "loss_gen")
    print("max |mean20(bf16) - mean20(fp32)| per curve:", {n: float(dev[:, i].max()) for i, n in enumerate(names)})
    for lo, hi in ((0, 30), (30, 80), (80, STEPS - WINDOW + 1)):
        print(f"  window starts {lo}..{hi}: max dev", [float(f"{x:.3f}") for x in dev[lo:hi].max(axis=0)], "fp32 mean", [float(f"{x:.3f}") for x in mr[lo:hi].mean(axis=0)])
    # controls: fp32 against fp32 with the initial weights perturbed - by 1e-6 relative (the chaotic spread of the GAN
    # itself) and by +-1e-3 relative (a perturbation of the size bf16 storage applies, half an ulp = 2e-3); the envelope
    # of their drifts is what any bf16 run has to be compared with once the trajectories have decorrelated
    cdevs = []
    for eps in (1e-6, 1e-3, -1e-3):
        init2 = {k: {a: (b * (1 + eps) if b.dtype == torch.float32 else b.clone()) for a, b in v.items()} for k, v in init.items()}
        cdevs.append(np.abs(_running_mean(_run(cfg, init2, "fp32", STEPS, tmp_path), WINDOW) - mr))
        print(f"control (fp32 vs {eps:+.0e}-perturbed fp32) max dev per curve:", [float(f"{x:.3f}") for x in cdevs[-1].max(axis=0)])
    cdev = np.maximum.reduce(cdevs)
    print("final means fp32:", mr[-1].tolist(), "bf16:", mg[-1].tolist())
    print("first-step losses fp32:", ref[0].tolist(), "bf16:", got[0].tolist())
    early_tol = np.maximum(0.05 * np.abs(mr[:30]), 0.05)
    assert (dev[:30] <= early_tol).all(), {n: float((dev[:30, i] / early_tol[:, i]).max()) for i, n in enumerate(names)}
    full_tol = np.maximum(2.0 * cdev.max(axis=0), 0.05)
    assert (dev.max(axis=0) <= full_tol).all(), dict(zip(names, (dev.max(axis=0) / full_tol).tolist()))


def test_fp32_curve_prefix_matches_oracle(tmp_path):
    """12 consecutive iterations of the fp32 CUDA path against the CPU oracle: float-rounding agreement on the first
    iteration (measured 3.6e-7), then the slow chaotic drift of two fp32 implementations (measured <= 0.05)."""
    cfg = small_cfg("depth", 1, "adversarial-loss", noise=False, ngf=8, ndf=8, gdis=False)
    init = orc.init_all(cfg, 10)
    o = orc.OracleTrainer(cfg, {k: {a: b.clone() for a, b in v.items()} for k, v in init.items()})
    torch.manual_seed(123)
    np.random.seed(123)
    ref = []
    for it in range(12):
        xc, xg = orc.synthetic_batch(cfg, cfg["batchsize"], 5000 + it % 8)
        r = o.step(xc, xg)
        ref.append([r["loss_idis"], r["loss_vdis"], 0.0, r["loss_gen"]])
    got = _run(cfg, init, "fp32", 12, tmp_path)
    dev = np.abs(got - np.array(ref)).max(axis=1)
    print("per-iteration max abs loss deviation:", [float(f"{d:.2e}") for d in dev])
    # rounding differences are amplified by the adversarial dynamics (Adam's first steps move every weight by
    # ~lr*sign(grad)); the deviation must start at float rounding level and stay small over the prefix
    assert dev[0] < 1e-5 and dev[1] < 2e-3 and dev.max() < 0.1
