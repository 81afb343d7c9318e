"""Drop-in replacements for the reference's src/loss.py (Loss, AdversarialLoss, HingeLoss).

Each term is one fused kernel launch (dcv_loss_fwd_bwd) that produces the mean loss and dL/dy together;
the returned 0-d tensors are differentiable, so `loss.backward()` works as in trainer.py:319,356.
"""
from abc import ABCMeta, abstractmethod

import torch

from . import _lib, ops
from .generator import current_device


class _LossTerm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, kind):
        _lib.require_device()
        if not y.is_cuda:
            raise _lib.DcvError("dcvgan_b200 losses only run on a B200 (no CPU fallback)")
        yc = y.detach().float().contiguous()
        out = torch.empty(1, dtype=torch.float32, device=y.device)
        dy = torch.empty_like(yc)
        ops.loss_fwd_bwd(yc, kind, out, False, dy, 1.0)
        ctx.save_for_backward(dy)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (dy,) = ctx.saved_tensors
        return dy * g, None


def loss_term(y, kind):
    return _LossTerm.apply(y, kind)


class Loss(object):
    """Abstract loss class (loss.py:9-58)."""

    __metaclass__ = ABCMeta

    @abstractmethod
    def compute_dis_loss(self, y_real: torch.Tensor, y_fake: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError()

    @abstractmethod
    def compute_gen_loss(self, y_fake_i: torch.Tensor, y_fake_v: torch.Tensor, y_fake_g: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError()


class AdversarialLoss(Loss):
    """BCE-with-logits against ones/zeros, summed then divided by numel (loss.py:61-131)."""

    name = "adversarial-loss"
    KIND_REAL, KIND_FAKE, KIND_GEN = _lib.LOSS_BCE_ONES, _lib.LOSS_BCE_ZEROS, _lib.LOSS_BCE_ONES
    gen_uses_gdis = True

    def __init__(self):
        super().__init__()
        self.device = current_device()

    def compute_dis_loss(self, y_real, y_fake):
        return loss_term(y_real, self.KIND_REAL) + loss_term(y_fake, self.KIND_FAKE)

    def compute_gen_loss(self, y_fake_i, y_fake_v, y_fake_g):
        return loss_term(y_fake_i, self.KIND_GEN) + loss_term(y_fake_v, self.KIND_GEN) + loss_term(y_fake_g, self.KIND_GEN)


class HingeLoss(Loss):
    """mean(relu(1 -/+ y)) for the discriminators; mean(softplus(-y)) of the image and video logits for the
    generators - the gradient discriminator's logits are ignored, as in loss.py:163-166,190-193."""

    name = "hinge-loss"
    KIND_REAL, KIND_FAKE, KIND_GEN = _lib.LOSS_HINGE_REAL, _lib.LOSS_HINGE_FAKE, _lib.LOSS_SOFTPLUS_NEG
    gen_uses_gdis = False

    def __init__(self):
        super().__init__()
        self.device = current_device()

    def compute_dis_loss(self, y_real, y_fake):
        return loss_term(y_real, self.KIND_REAL) + loss_term(y_fake, self.KIND_FAKE)

    def compute_gen_loss(self, y_fake_i, y_fake_v, y_fake_g):
        return loss_term(y_fake_i, self.KIND_GEN) + loss_term(y_fake_v, self.KIND_GEN)
