"""Loading of the golden fixtures produced by oracle/make_golden.py."""
import json
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["flow_hinge_noise", "depth_adv_nogdis", "segm_adv_noise", "depth_hinge_evalquirk"]
LOG2 = 0.6931471805599453


def load_case(name):
    meta = json.loads((GOLDEN / f"{name}.json").read_text())
    arrays = np.load(GOLDEN / f"{name}.npz")
    init = {}
    for key in arrays.files:
        _, net, k = key.split("/", 2)
        init.setdefault(net, {})[k] = torch.from_numpy(arrays[key]).clone()
    return meta, init


def expected_losses(meta, it):
    """Reference-logged losses of iteration `it`, with the constants the 'gdis disabled' stub contributed removed."""
    l = dict(meta["losses"][it])
    if not meta["cfg"]["gdis"]["enabled"]:
        l.pop("loss_gdis")
        if meta["cfg"]["loss"] == "adversarial-loss":
            l["loss_gen"] -= LOG2      # BCEWithLogits(0, 1) of the stub's all-zero logits
    return l


def _is_buffer(k):
    return k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")


def check_digest(state_dict, dig, rtol, atol, what="", lr_steps=None):
    """Compare a state_dict with a recorded digest (32 sampled entries + sum per tensor).

    BatchNorm buffers are compared strictly (rtol/atol).  Parameters have been through Adam, whose first
    steps move every weight by ~lr * sign(grad): an entry whose gradient is zero to rounding may land one
    full step away from the reference.  With `lr_steps` (= lr * number of Adam steps taken) given, such
    entries are tolerated when they stay within 2.5 * lr_steps and are fewer than 15 % of the samples.
    """
    worst = 0.0
    for k, d in dig.items():
        v = state_dict[k].detach().double().flatten().cpu()
        got = v[torch.tensor(d["idx"])]
        exp = torch.tensor(d["val"], dtype=torch.float64)
        ratio = (got - exp).abs() / (atol + rtol * exp.abs())
        if lr_steps is not None and not _is_buffer(k):
            bad = ratio > 1.0
            assert float(((got - exp).abs()[bad]).max() if bad.any() else 0.0) <= 2.5 * lr_steps, f"{what}{k}: entry further than 2.5 Adam steps away"
            assert float(bad.double().mean()) <= 0.15, f"{what}{k}: {int(bad.sum())}/32 sampled entries off by a sign-flipped Adam step"
            continue
        err = float(ratio.max())
        worst = max(worst, err)
        assert err <= 1.0, f"{what}{k}: sampled values differ (worst ratio {err:.3g}): got {got[:4].tolist()} expected {exp[:4].tolist()}"
        n = v.numel()
        assert abs(float(v.sum()) - d["sum"]) <= n * atol + rtol * d["abs"] + 1e-12, f"{what}{k}: sum differs"
    return worst
