"""Generate tests/golden/curves_flow_hinge.json: 200-iteration loss curves of the UNMODIFIED reference (modules +
Trainer.train(), driven by tools/ref_harness.py) on CPU for several seeds - the denominators of the bf16 loss-curve gate
(tests/test_curves_gpu.py).  Run in the build container, where /root/reference is mounted (takes ~10 minutes):

    python oracle/make_curves.py

Per seed s: initial weights orc.init_all(cfg, 100 + s) loaded into the reference modules through state_dict, torch /
numpy seeded with 1000 + s right before Trainer.train(), batches synthetic_batch(cfg, B, 5000 + it % 8).  The GPU test
replays exactly this with the CUDA path in 'cpu_parity' RNG mode, so both sides consume bit-identical latents, Noise and
Dropout draws.
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import dcvgan_oracle as orc  # noqa: E402
from tools import ref_harness as rh  # noqa: E402

STEPS, SEEDS = 200, tuple(range(12))
O = {"lr": 0.0002, "decay": 0.00001}
CFG = {"batchsize": 4, "video_length": 16, "image_size": 64, "geometric_info": {"name": "optical-flow", "channel": 2},
       "loss": "hinge-loss", "num_gen_update": 1, "num_dis_update": 1, "n_epochs": 1, "seed": 0,
       "log_interval": 10 ** 9, "log_samples_interval": 10 ** 9, "snapshot_interval": 10 ** 9, "evaluation_interval": 10 ** 9,
       "evaluation": {"batchsize": 2, "num_samples": 0, "metrics": []},
       "ggen": {"dim_z_content": 40, "dim_z_motion": 10, "ngf": 16, "optimizer": O},
       "cgen": {"dim_z_color": 10, "ngf": 16, "optimizer": O},
       "idis": {"use_noise": True, "noise_sigma": 0.2, "ndf": 16, "optimizer": O},
       "vdis": {"use_noise": True, "noise_sigma": 0.2, "ndf": 16, "optimizer": O},
       "gdis": {"use_noise": True, "noise_sigma": 0.2, "ndf": 16, "optimizer": O, "enabled": True}}


def run_seed(ref, s):
    cfg = CFG
    gname = cfg["geometric_info"]["name"]
    models, optimizers, loss = rh.build_reference_models(ref, cfg, "cpu")
    init = orc.init_all(cfg, 100 + s)
    for k, m in models.items():
        m.load_state_dict({a: b.clone() for a, b in init[k].items()})
    pool = []
    for i in range(8):
        xc, xg = orc.synthetic_batch(cfg, cfg["batchsize"], 5000 + i)
        pool.append({"color": xc, gname: xg})
    batches = [pool[it % 8] for it in range(STEPS)]
    _, logger = rh.run_reference_trainer(ref, cfg, batches, device="cpu", seed=1000 + s, models=(models, optimizers, loss))
    L = logger.losses(STEPS)
    return [[l["loss_idis"], l["loss_vdis"], l["loss_gdis"], l["loss_gen"]] for l in L]


if __name__ == "__main__":
    torch.set_num_threads(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
    ref = rh.import_reference()
    path = ROOT / "tests" / "golden" / "curves_flow_hinge.json"
    out = {"cfg": CFG, "steps": STEPS, "torch": torch.__version__, "seeds": {}}
    if path.exists():                    # seeds are independent runs: keep the ones already recorded
        old = json.loads(path.read_text())
        if old["cfg"] == CFG and old["steps"] == STEPS:
            out["seeds"] = old["seeds"]
    for s in SEEDS:
        if str(s) in out["seeds"]:
            continue
        out["seeds"][str(s)] = run_seed(ref, s)
        print("seed", s, "first", out["seeds"][str(s)][0], "last", out["seeds"][str(s)][-1], flush=True)
        path.write_text(json.dumps(out))
