"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container, where /root/reference is mounted:
    python oracle/make_golden.py
It imports the reference's own modules (util first - import cycle), builds the
five models exactly as train.py:117-176 does, and drives the unmodified `Trainer.train()` loop
(trainer.py:226-392) on a synthetic in-memory "dataloader".  Only the side work that is not on the
hot path is neutralised (log_samples / evaluate / save_params / log_hparams and the third-party
imports that are not installed here).  Nothing is written into /root/reference.

Each fixture holds: the normalised config, the initial state_dicts, the seeds (and digests) of the real batches, the RNG seed,
the per-iteration losses the reference logged and a digest (sum, abs-sum, 32 samples per tensor) of
every state_dict entry after training.  tests/test_oracle.py replays them through
oracle/dcvgan_oracle.py (bit-level agreement expected on the same torch build); tests/test_step_gpu.py
replays them through the CUDA path.
"""
import copy
import json
import logging
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/src")
OUT = ROOT / "tests" / "golden"

sys.path.insert(0, str(ROOT))
from oracle import dcvgan_oracle as orc  # noqa: E402  (only for normalise_config / synthetic_batch)


def import_reference():
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(REF))
    for n in ["skvideo", "skvideo.io", "evan", "colorlog", "tensorboardX", "matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["skvideo"].io = sys.modules["skvideo.io"]

    class _SW:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return lambda *a, **k: None

    sys.modules["tensorboardX"].SummaryWriter = _SW
    sys.modules["colorlog"].ColoredFormatter = lambda fmt, datefmt=None: logging.Formatter(fmt.replace("%(log_color)s", ""), datefmt)
    import util  # noqa: F401  (must come first)
    import generator, discriminator, loss, trainer  # noqa: E401
    trainer.Trainer.log_samples = lambda self, *a, **k: None
    trainer.Trainer.evaluate = lambda self, *a, **k: None
    trainer.Trainer.save_params = lambda self, *a, **k: None
    trainer.Trainer.log_hparams = lambda self, *a, **k: None
    trainer.Trainer.save_classobj = lambda self, *a, **k: None
    trainer.VideoDataLoader = lambda *a, **k: None
    return util, generator, discriminator, loss, trainer


class ListLoader:
    """Stands in for VideoDataLoader: yields prebuilt batches, consumes no RNG."""

    def __init__(self, batches):
        self.batches = batches
        self.dataset = types.SimpleNamespace(root_path=Path("/tmp"))

    def __iter__(self):
        return iter(self.batches)


class CaptureLogger:
    def __init__(self, path):
        self.path = Path(path)
        self.records = []

    def update(self, name, value):
        self.records.append((name, value))

    def __getattr__(self, name):
        return lambda *a, **k: None


class StubGdis(torch.nn.Module):
    """'gdis disabled' stand-in: contributes constants to the logged losses, no gradient, no RNG draw."""

    def __init__(self):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, xg, xc):
        return torch.zeros((xg.shape[0], 3, 4, 4))


def base_cfg(name, channel, loss, **over):
    cfg = {
        "experiment_name": "golden", "batchsize": 2, "n_epochs": 1, "seed": 7, "video_length": 16, "image_size": 64,
        "log_dir": "/tmp", "tensorboard_dir": "/tmp", "geometric_info": {"name": name, "channel": channel},
        "log_interval": 10 ** 9, "log_samples_interval": 10 ** 9, "snapshot_interval": 10 ** 9, "evaluation_interval": 10 ** 9,
        "loss": loss, "num_gen_update": 1, "num_dis_update": 1,
        "dataset": {"name": "synthetic", "path": "/tmp", "n_workers": 0, "number_limit": -1},
        "evaluation": {"batchsize": 2, "num_samples": 0, "metrics": []},
        "ggen": {"dim_z_content": 40, "dim_z_motion": 10, "ngf": 6, "optimizer": {"lr": 0.0002, "decay": 0.00001}},
        "cgen": {"dim_z_color": 10, "ngf": 6, "optimizer": {"lr": 0.0002, "decay": 0.00001}},
        "idis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 8, "optimizer": {"lr": 0.0002, "decay": 0.00001}},
        "vdis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 8, "optimizer": {"lr": 0.0002, "decay": 0.00001}},
        "gdis": {"use_noise": False, "noise_sigma": 0.2, "ndf": 8, "optimizer": {"lr": 0.0002, "decay": 0.00001}, "enabled": True},
    }
    for k, v in over.items():
        if isinstance(v, dict):
            cfg[k].update(v)
        else:
            cfg[k] = v
    return cfg


CASES = {
    # isogd-flow-like: hinge loss, Noise on idis/vdis, gdis without noise
    "flow_hinge_noise": (base_cfg("optical-flow", 2, "hinge-loss", idis={"use_noise": True}, vdis={"use_noise": True}), 2),
    # mug/surreal-depth-like: adversarial loss, gdis disabled, D updated every 2nd iteration
    "depth_adv_nogdis": (base_cfg("depth", 1, "adversarial-loss", num_gen_update=2, gdis={"enabled": False}), 2),
    # surreal-segm-like: 25-channel softmax geometry, adversarial loss, Noise everywhere incl. gdis
    "segm_adv_noise": (base_cfg("segmentation", 25, "adversarial-loss", idis={"use_noise": True}, vdis={"use_noise": True},
                                gdis={"use_noise": True}), 1),
    # trainer.py:126-127 quirk: log_samples() leaves both generators in eval mode, so the D-phase fakes of the NEXT iteration
    # are generated with running-statistics BatchNorm and without Dropout (and loss_dis.backward() runs through eval-mode
    # generators).  Here log_samples is reduced to exactly that side effect (ggen.eval(); cgen.eval(), no RNG draw) and fires
    # before iteration 1 and after iteration 2 (log_samples_interval 2), so iterations 1 and 3 see eval-mode fakes.
    "depth_hinge_evalquirk": (base_cfg("depth", 1, "hinge-loss", log_samples_interval=2, eval_quirk=True, gdis={"enabled": False}), 4),
}


def digest(sd, gen):
    out = {}
    for k, v in sd.items():
        v = v.detach().double().flatten()
        idx = torch.randint(0, v.numel(), (32,), generator=gen)
        out[k] = {"sum": float(v.sum()), "abs": float(v.abs().sum()), "idx": idx.tolist(), "val": v[idx].tolist()}
    return out


def run_case(name, cfg, iters, ref):
    util, generator, discriminator, loss_mod, trainer = ref
    torch.manual_seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    C, gname = cfg["geometric_info"]["channel"], cfg["geometric_info"]["name"]
    ggen = generator.GeometricVideoGenerator(cfg["ggen"]["dim_z_content"], cfg["ggen"]["dim_z_motion"], C, gname,
                                             cfg["ggen"]["ngf"], cfg["video_length"])
    cgen = generator.ColorVideoGenerator(ggen.channel, cfg["cgen"]["dim_z_color"], gname, cfg["cgen"]["ngf"], cfg["video_length"])
    idis = discriminator.ImageDiscriminator(C, 3, cfg["idis"]["use_noise"], cfg["idis"]["noise_sigma"], cfg["idis"]["ndf"])
    vdis = discriminator.VideoDiscriminator(C, 3, cfg["vdis"]["use_noise"], cfg["vdis"]["noise_sigma"], cfg["vdis"]["ndf"])
    if cfg["gdis"]["enabled"]:
        gdis = discriminator.GradientDiscriminator(C, 3, cfg["gdis"]["use_noise"], cfg["gdis"]["noise_sigma"], cfg["gdis"]["ndf"])
    else:
        gdis = StubGdis()
    models = {"ggen": ggen, "cgen": cgen, "idis": idis, "vdis": vdis, "gdis": gdis}
    for m in models.values():
        m.apply(util.init_weights)                                               # train.py:164-165
    optimizers = {}
    for k, m in models.items():                                                  # train.py:167-176
        o = cfg[k]["optimizer"]
        optimizers[k] = torch.optim.Adam(m.parameters(), lr=o["lr"], betas=(0.5, 0.999), weight_decay=o["decay"])
    init_sd = {k: copy.deepcopy(m.state_dict()) for k, m in models.items() if not isinstance(m, StubGdis)}
    batches = []
    for i in range(iters):
        xc, xg = orc.synthetic_batch(cfg, cfg["batchsize"], 1000 + i)
        batches.append({"color": xc, gname: xg})
    loss = loss_mod.AdversarialLoss() if cfg["loss"] == "adversarial-loss" else loss_mod.HingeLoss()
    tmp = tempfile.mkdtemp()
    cfg_path = Path(tmp) / "cfg.yml"
    cfg_path.write_text("golden: true\n")
    run_cfg = dict(cfg, config_path=str(cfg_path))
    logger = CaptureLogger(tmp)
    tr = trainer.Trainer(ListLoader(batches), logger, models, optimizers, loss, run_cfg)
    if cfg.get("eval_quirk"):
        def _log_samples(ggen_, cgen_, iteration):       # trainer.py:126-127, nothing else of log_samples
            ggen_.eval()
            cgen_.eval()
        tr.log_samples = _log_samples
    step_seed = cfg["seed"] + 100
    torch.manual_seed(step_seed)
    np.random.seed(step_seed)
    tr.train()
    losses = [{} for _ in range(iters)]
    it = -1
    for k, v in logger.records:
        if k == "iteration":
            it = v - 1
        elif k.startswith("loss_"):
            losses[it][k] = v
    gen = torch.Generator().manual_seed(0)
    final = {k: digest(m.state_dict(), gen) for k, m in models.items() if not isinstance(m, StubGdis)}
    arrays = {}
    for net, sd in init_sd.items():
        for k, v in sd.items():
            arrays[f"init/{net}/{k}"] = v.numpy()
    # the real batches are regenerated from their seeds (oracle.synthetic_batch); keep a digest to detect RNG drift
    batch_digest = [{"seed": 1000 + i, "color_sum": float(b["color"].double().sum()), "geo_sum": float(b[gname].double().sum()),
                     "color_first": b["color"].flatten()[:4].tolist()} for i, b in enumerate(batches)]
    meta = {"cfg": cfg, "iters": iters, "step_seed": step_seed, "losses": losses, "final": final,
            "torch": torch.__version__, "batches": batch_digest}
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"{name}.npz", **arrays)
    (OUT / f"{name}.json").write_text(json.dumps(meta))
    print(name, "losses", losses, "npz MB", round((OUT / f"{name}.npz").stat().st_size / 2 ** 20, 2))


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = import_reference()
    only = sys.argv[1:]
    for name, (cfg, iters) in CASES.items():
        if only and name not in only:
            continue
        run_case(name, cfg, iters, ref)
