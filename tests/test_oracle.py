"""Pins oracle/dcvgan_oracle.py against the golden vectors recorded from the unmodified reference
(modules + Trainer.train loop).  CPU only; same torch build => agreement to float rounding."""
import numpy as np
import pytest
import torch

from golden_util import CASES, check_digest, expected_losses, load_case
from oracle import dcvgan_oracle as orc


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_training(name):
    meta, init = load_case(name)
    cfg = meta["cfg"]
    torch.set_num_threads(8)
    if "gdis" not in init:
        init["gdis"] = {}
    tr = orc.OracleTrainer(cfg, init)
    torch.manual_seed(meta["step_seed"])
    np.random.seed(meta["step_seed"])
    quirk = cfg.get("eval_quirk", False)      # log_samples() before the loop and every log_samples_interval iterations
    if quirk:
        tr.gen_training = False               # trainer.py:126-127 via trainer.py:268
    for it in range(meta["iters"]):
        bd = meta["batches"][it]
        xc, xg = orc.synthetic_batch(cfg, cfg["batchsize"], bd["seed"])
        assert abs(float(xc.double().sum()) - bd["color_sum"]) < 1e-6 and abs(float(xg.double().sum()) - bd["geo_sum"]) < 1e-6
        out = tr.step(xc, xg)
        for k, v in expected_losses(meta, it).items():
            assert abs(out[k] - v) <= 2e-6 * max(1.0, abs(v)), (name, it, k, out[k], v)
        if quirk and (it + 1) % cfg["log_samples_interval"] == 0:
            tr.gen_training = False           # trainer.py:375-376
    for net, dig in meta["final"].items():
        check_digest(tr.P[net], dig, rtol=1e-5, atol=1e-7, what=f"{name}/{net}/")


def test_config_normaliser_on_legacy_schema():
    legacy = {"batchsize": 20, "gen": {"dim_z_content": 40, "dim_z_motion": 10, "dim_z_color": 10, "ngf": 64,
                                       "optimizer": {"lr": 0.0002, "decay": 0.00001}},
              "idis": {"use_noise": False, "noise_sigma": 0.1, "ndf": 64, "optimizer": {"lr": 0.0002, "decay": 0.00001}},
              "vdis": {"use_noise": False, "noise_sigma": 0.1, "ndf": 64, "optimizer": {"lr": 0.0002, "decay": 0.00001}}}
    cfg = orc.normalise_config(legacy)
    assert cfg["geometric_info"] == {"name": "depth", "channel": 1}
    assert cfg["ggen"]["ngf"] == 64 and cfg["cgen"]["dim_z_color"] == 10
    assert cfg["loss"] == "adversarial-loss" and cfg["num_gen_update"] == 1 and cfg["gdis"]["enabled"] is False


def test_gru_restatement_matches_torch():
    torch.manual_seed(0)
    cell = torch.nn.GRUCell(10, 10)
    x, h = torch.randn(4, 10), torch.randn(4, 10)
    ref = cell(x, h)
    mine = orc.gru_cell(x, h, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
    assert float((ref - mine).abs().max()) < 1e-6
