"""Shared helpers for the GPU parity tests."""
import torch

from dcvgan_b200 import ops
from dcvgan_b200.ops import Act


def to_act(x, dtype, ld=None):
    """x: CPU/GPU fp32 (N,C,H,W) or (N,C,T,H,W) -> channels-last Act on the GPU."""
    x = x.cuda().float()
    if x.dim() == 4:
        x = x.unsqueeze(2)
    n, c, t, h, w = x.shape
    a = Act.empty(n, t, h, w, c, dtype, ld=ld)
    ops.to_channels_last(x, a)
    return a


def from_act(a, dims=5):
    """channels-last Act -> CPU fp32 (N,C,T,H,W) (or (N,C,H,W) when dims == 4)"""
    out = torch.empty((a.n, a.c, a.t, a.h, a.w), dtype=torch.float32, device="cuda")
    ops.from_channels_last(a, out)
    torch.cuda.synchronize()
    out = out.cpu()
    return out[:, :, 0] if dims == 4 else out


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cos_sim(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def trainer_without_pickles(trainer_mod, *args):
    """Trainer(*args) without the whole-module pickles of save_classobj (trainer.py:70-76: slow and irrelevant to these tests).
    The class attribute is restored before returning, so later tests in the same process (the drop-in run of train.py checks
    those very files) see the real method."""
    orig = trainer_mod.Trainer.save_classobj
    trainer_mod.Trainer.save_classobj = lambda self: None
    try:
        return trainer_mod.Trainer(*args)
    finally:
        trainer_mod.Trainer.save_classobj = orig
