"""Harness that drives the UNMODIFIED reference (raahii/dcvgan `src/`) for baselines and golden vectors.

The reference is a plain Python source tree without packaging metadata, so "installing" it means staging its `src/`
directory: `stage_reference()` copies /root/reference/src/*.py into the git-ignored `baseline/_ref/src/` (done by
`__graft_entry__.build()` in the build container; the directory travels to the GPU box with the gpurun snapshot, the
read-only /root/reference does not).  Nothing here is imported by the dcvgan_b200 package: callers are
`bench.py` (reference arm, cpu_baseline and gpu_eager_baseline legs), `oracle/make_golden.py` and `tests/`.

`import_reference()` makes the staged modules importable (util first - import cycle util.py:13 <-> generator.py:8) with
empty stand-ins for the third-party packages that are not installed (skvideo, evan, colorlog, tensorboardX,
matplotlib) and neutralises only the side work that is not on the training-step path (log_samples / evaluate /
save_params / log_hparams / save_classobj).  `run_reference_trainer()` builds the five models exactly as
train.py:117-176 does and runs the unmodified `Trainer.train()` (trainer.py:226-392) over in-memory batches.
"""
import logging
import os
import shutil
import sys
import tempfile
import time
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF_ORIGIN = Path("/root/reference/src")
REF_STAGED = ROOT / "baseline" / "_ref" / "src"


def stage_reference():
    """copy the reference's src/*.py into baseline/_ref/src (git-ignored); returns the staged path or None"""
    if not REF_ORIGIN.exists():
        return REF_STAGED if REF_STAGED.exists() else None
    for f in list(REF_ORIGIN.glob("*.py")) + list((REF_ORIGIN / "preprocess").glob("*.py")):
        dst = REF_STAGED / f.relative_to(REF_ORIGIN)
        dst.parent.mkdir(parents=True, exist_ok=True)
        if not dst.exists() or dst.read_bytes() != f.read_bytes():
            shutil.copyfile(f, dst)
    return REF_STAGED


def reference_src():
    """where the reference can be imported from in this process: the staged copy, else the container's mount, else None"""
    if REF_STAGED.exists() and (REF_STAGED / "trainer.py").exists():
        return REF_STAGED
    if REF_ORIGIN.exists():
        return REF_ORIGIN
    return None


_REF = None


def import_reference(src=None):
    global _REF
    if _REF is not None:
        return _REF
    src = Path(src) if src is not None else reference_src()
    if src is None:
        raise RuntimeError("reference sources are not staged (baseline/_ref/src) and /root/reference is absent")
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(src))
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
    _REF = (util, generator, discriminator, loss, trainer)
    return _REF


class ListLoader:
    """Stands in for VideoDataLoader: yields prebuilt batches, consumes no RNG."""

    def __init__(self, batches):
        self.batches = batches
        self.dataset = types.SimpleNamespace(root_path=Path("/tmp"))

    def __iter__(self):
        return iter(self.batches)


class CaptureLogger:
    """records every logger.update(); `sync` (a callable) runs before each timestamp so GPU runs are timed honestly"""

    def __init__(self, path, sync=None):
        self.path = Path(path)
        self.records = []
        self.iter_start = []
        self.sync = sync

    def update(self, name, value):
        if name == "iteration":
            if self.sync is not None:
                self.sync()
            self.iter_start.append(time.perf_counter())
        self.records.append((name, value))

    def losses(self, iters):
        out = [{} for _ in range(iters)]
        it = -1
        for k, v in self.records:
            if k == "iteration":
                it = v - 1
            elif k.startswith("loss_"):
                out[it][k] = v
        return out

    def __getattr__(self, name):
        return lambda *a, **k: None


class StubGdis(torch.nn.Module):
    """'gdis disabled' stand-in (configs without a gradient discriminator): contributes constants to the logged losses,
    no gradient, no RNG draw, no measurable time."""

    def __init__(self):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, xg, xc):
        return torch.zeros((xg.shape[0], 3, 4, 4), device=xg.device)


def build_reference_models(ref, cfg, device):
    """the five models and optimizers as train.py:117-176 builds them, living on `device`"""
    util, generator, discriminator, loss_mod, trainer = ref
    saved = util.current_device
    util.current_device = lambda: torch.device(device)          # the modules capture their device at construction
    try:
        C, gname = cfg["geometric_info"]["channel"], cfg["geometric_info"]["name"]
        ggen = generator.GeometricVideoGenerator(cfg["ggen"]["dim_z_content"], cfg["ggen"]["dim_z_motion"], C, gname,
                                                 cfg["ggen"]["ngf"], cfg["video_length"])
        cgen = generator.ColorVideoGenerator(ggen.channel, cfg["cgen"]["dim_z_color"], gname, cfg["cgen"]["ngf"], cfg["video_length"])
        idis = discriminator.ImageDiscriminator(C, 3, cfg["idis"]["use_noise"], cfg["idis"]["noise_sigma"], cfg["idis"]["ndf"])
        vdis = discriminator.VideoDiscriminator(C, 3, cfg["vdis"]["use_noise"], cfg["vdis"]["noise_sigma"], cfg["vdis"]["ndf"])
        if cfg["gdis"].get("enabled", True):
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
        loss = loss_mod.AdversarialLoss() if cfg["loss"] == "adversarial-loss" else loss_mod.HingeLoss()
    finally:
        util.current_device = saved
    return models, optimizers, loss


def run_reference_trainer(ref, cfg, batches, device="cpu", seed=None, autocast_dtype=None, channels_last=False, models=None):
    """Unmodified Trainer.train() over `batches` (list of {"color": ..., <geometry name>: ...}) on `device`.
    Returns (models, logger): logger.iter_start holds a perf_counter() timestamp per iteration start (after a device
    sync when device is CUDA), logger.t_end the one after the loop."""
    util, generator, discriminator, loss_mod, trainer = ref
    if models is None:
        models, optimizers, loss = build_reference_models(ref, cfg, device)
    else:
        models, optimizers, loss = models
    if channels_last:
        for m in models.values():
            for p_ in m.modules():
                if isinstance(p_, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                    p_.to(memory_format=torch.channels_last)
                elif isinstance(p_, torch.nn.Conv3d):
                    p_.to(memory_format=torch.channels_last_3d)
    tmp = tempfile.mkdtemp()
    cfg_path = Path(tmp) / "cfg.yml"
    cfg_path.write_text("synthetic: true\n")
    run_cfg = dict(cfg, config_path=str(cfg_path))
    is_cuda = str(device).startswith("cuda")
    logger = CaptureLogger(tmp, sync=torch.cuda.synchronize if is_cuda else None)
    saved = util.current_device
    util.current_device = lambda: torch.device(device)
    try:
        tr = trainer.Trainer(ListLoader(batches), logger, models, optimizers, loss, run_cfg)
        if seed is not None:
            torch.manual_seed(seed)
            np.random.seed(seed)
        if autocast_dtype is not None:
            with torch.autocast("cuda" if is_cuda else "cpu", dtype=autocast_dtype):
                tr.train()
        else:
            tr.train()
    finally:
        util.current_device = saved
    if is_cuda:
        torch.cuda.synchronize()
    logger.t_end = time.perf_counter()
    return models, logger


def timed_iters_per_s(logger, warmup, steps):
    """iterations/s over exactly `steps` iterations after `warmup` (timestamps taken at iteration starts + loop end)"""
    t = logger.iter_start + [logger.t_end]
    assert len(t) >= warmup + steps + 1, (len(t), warmup, steps)
    dt = t[warmup + steps] - t[warmup]
    return steps / dt, dt / steps
