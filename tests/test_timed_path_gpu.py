"""Parity of the path bench.py times: bf16 / tcgen05, CUDA-graph replay, batch 32, config/mug-depth.yml shapes.

The graph path cannot take host-generated noise inside the iteration, so these tests use the 'staged' RNG mode
(engine.Rng): the reference's CPU draws (same order, same torch generator) are copied into static device buffers
BEFORE the iteration, and the captured graph reads them - the replayed iteration therefore consumes exactly the
latents / Dropout masks the CPU oracle consumes.
"""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import cos_sim, rel_err, trainer_without_pickles  # noqa: E402
from oracle import dcvgan_oracle as orc  # noqa: E402
from test_nets_gpu import _Logger, _mods, _net_cosines, build_models, small_cfg  # noqa: E402


def _mug_depth_cfg(B, ngf=64, ndf=64):
    cfg = small_cfg("depth", 1, "adversarial-loss", noise=False, ngf=ngf, ndf=ndf, gdis=False)
    cfg["batchsize"] = B
    cfg["gdis"]["ndf"] = 32
    return cfg


def _trainer(cfg, init, precision, tmp_path, rng_mode):
    dcv, _, _, loss_mod, trainer_mod, _, engine = _mods()
    init = {k: v for k, v in init.items() if k != "gdis" or cfg["gdis"].get("enabled", True)}
    models = build_models(cfg, init, precision)
    engine.set_rng_mode(rng_mode)
    opts = {k: torch.optim.Adam(m.parameters(), lr=cfg[k]["optimizer"]["lr"], betas=(0.5, 0.999),
                                weight_decay=cfg[k]["optimizer"]["decay"]) for k, m in models.items()}
    L = loss_mod.AdversarialLoss() if cfg["loss"] == "adversarial-loss" else loss_mod.HingeLoss()
    return trainer_without_pickles(trainer_mod, None, _Logger(tmp_path), models, opts, L, dict(cfg, config_path="")), models, engine


def test_graph_replay_batch32_bf16_matches_oracle(tmp_path):
    """ONE iteration of the benchmarked configuration (mug-depth widths, batch 32, bf16, CUDA-graph REPLAY) against the CPU
    oracle from identical weights, batch, frame index and noise.  Gates: losses within 2 %, per-network gradient cosine
    >= 0.99 (concatenated parameter gradients of each network), BatchNorm running statistics within 1e-2 (norm-relative)."""
    B = 32
    cfg = _mug_depth_cfg(B)
    init = orc.init_all(cfg, 41)
    # a throw-away trainer performs the lazy one-off work (module loads, shared-memory opt-ins) and records the staged RNG
    # request pattern, so that the trainer under test can capture its very first iteration
    warm, _, engine = _trainer(cfg, copy.deepcopy(init), "bf16", tmp_path, "staged")
    rng = engine.rng()
    xc, xg = orc.synthetic_batch(cfg, B, 3001)
    xc_d, xg_d = xc.cuda(), xg.cuda()
    rng.stage()
    warm.iteration = 1
    warm.use_cuda_graph = False
    warm.train_step(xc_d, xg_d, t_rand=5)
    del warm
    torch.cuda.empty_cache()

    tr, models, _ = _trainer(cfg, copy.deepcopy(init), "bf16", tmp_path, "staged")
    engine._RNG = rng                     # keep the request pattern recorded by the warm-up trainer
    tr.GRAPH_WARMUP = 0
    assert tr.use_cuda_graph
    # oracle (CPU, fp32): same seed -> same draws as the staged buffers below
    torch.set_num_threads(os.cpu_count())
    o = orc.OracleTrainer(cfg, {k: {a: b.clone() for a, b in v.items()} for k, v in init.items()})
    o.capture_grads = True
    torch.manual_seed(77)
    np.random.seed(77)
    ref = o.step(xc, xg, t_rand=5)
    torch.manual_seed(77)
    np.random.seed(77)
    rng.stage()
    tr.iteration = 1
    got = tr.train_step(xc_d, xg_d, t_rand=5).cpu().tolist()
    torch.cuda.synchronize()
    slot = next(iter(tr._graphs.values()))
    assert slot[1] is not None and tr.replayed_launches > 0, "the iteration under test must have run as a graph replay"
    engine.set_rng_mode("cpu_parity")
    names = ("loss_idis", "loss_vdis", "loss_gdis", "loss_gen")
    dev = {}
    for n, v in zip(names, got):
        if ref.get(n) is None:
            continue
        dev[n] = abs(v - ref[n]) / max(1.0, abs(ref[n]))
    my_g = {n: {k: p.grad.detach().cpu().clone() for k, p in models[n].named_parameters()} for n in models}
    nets = _net_cosines(my_g, dict(o.g_grads, **o.d_grads))
    worst_bn = 0.0
    for net in models:
        sd = models[net].state_dict()
        for k, v in o.P[net].items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                worst_bn = max(worst_bn, rel_err(sd[k].cpu(), v.detach()))
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v), (net, k)
    print(f"graph replay B=32 bf16 vs oracle: losses {dict(zip(names, got))} ref {ref}; relative loss deviation {dev}; "
          f"per-network gradient cosine {nets}; worst BatchNorm running-stat error {worst_bn:.2e}")
    assert max(dev.values()) <= 2e-2, dev
    assert min(nets.values()) >= 0.99, nets
    assert worst_bn <= 1e-2, worst_bn


def test_eager_steps_after_graph_replays_use_fresh_weights(tmp_path):
    """graph replays run Adam on the device; an eager iteration of a NOT YET captured update pattern
    that follows must not reuse packed bf16 weights left in the host-side cache by an earlier eager iteration.
    Interleaves two graph keys (generators in train mode / left in eval mode by log_samples) - eager, eager, capture+replay,
    replay, then the first eager iteration of the second key - and compares every loss and the final weights with the same
    schedule run without CUDA graphs."""
    cfg = _mug_depth_cfg(2, ngf=16, ndf=16)
    init = orc.init_all(cfg, 17)
    xc, xg = orc.synthetic_batch(cfg, 2, 9)
    xc_d, xg_d = xc.cuda(), xg.cuda()
    runs = {}
    for use_graph in (True, False):
        tr, models, engine = _trainer(cfg, copy.deepcopy(init), "bf16", tmp_path, "staged")
        tr.use_cuda_graph = use_graph
        rng_train, rng_eval = engine.Rng("staged"), engine.Rng("staged")      # one request pattern per update pattern
        torch.manual_seed(5)
        losses = []
        for it in range(1, 8):
            tr.iteration = it
            if it in (5, 7):                       # trainer.py:126-127: log_samples() leaves the generators in eval mode
                tr.log_samples(models["ggen"], models["cgen"], it)
            engine._RNG = rng_eval if it in (5, 7) else rng_train
            engine._RNG.stage()
            losses.append(tr.train_step(xc_d, xg_d, t_rand=it % 16).cpu())
        torch.cuda.synchronize()
        if use_graph:
            assert any(s[1] is not None for s in tr._graphs.values()), "train-mode pattern should have been captured"
        runs[use_graph] = (torch.stack(losses), {n: {k: p.detach().cpu().clone() for k, p in m.named_parameters()} for n, m in models.items()})
        engine.set_rng_mode("cpu_parity")
    Lg, Pg = runs[True]
    Le, Pe = runs[False]
    print("graph vs eager losses max abs diff per iteration:", (Lg - Le).abs().max(dim=1).values.tolist())
    assert torch.allclose(Lg, Le, rtol=0, atol=1e-5), (Lg, Le)
    for n in Pg:
        for k in Pg[n]:
            assert torch.equal(Pg[n][k], Pe[n][k]), (n, k)
