"""The plain-C oracle (CPU baseline of bench.py) against the torch-autograd oracle."""

import numpy as np
import torch

from helpers import assert_close
from oracle import c_oracle
from oracle import models_ref as mr


def _case(lorentzian, seed):
    rng = np.random.default_rng(seed)
    B, N, S_e, S_a = 12, 3, 150, 90
    ve, va = np.linspace(-50, 50, S_e), np.linspace(-25, 25, S_a)
    det = {
        "velocity": rng.normal(0, 8, (B, N)), "fwhm2": rng.uniform(5, 200, (B, N)),
        "fwhm_L": rng.uniform(0.1, 3.0, (B, N)) if lorentzian else np.zeros((B, N)),
        "tau_total": rng.uniform(0.05, 8, (B, N)), "tspin": rng.uniform(30, 3000, (B, N)),
        "filling_factor": rng.uniform(0.1, 1.0, (B, N)), "absorption_weight": rng.uniform(0.3, 1.2, (B, N)),
    }
    data = {"emission": mr.OracleData(ve, rng.normal(5, 3, S_e), 0.3),
            "absorption": mr.OracleData(va, rng.normal(0.2, 0.1, S_a), 0.02)}
    return det, data


def test_c_oracle_matches_autograd():
    spec = mr.make_spec("emission_absorption", 3, bg_temp=3.77)
    for lor, seed in ((False, 1), (True, 2)):
        det, data = _case(lor, seed)
        t = {k: torch.tensor(v, requires_grad=True) for k, v in det.items()}
        lp = mr.logp_from_inputs(spec, data, t)
        lp.sum().backward()
        c_lp, c_g, used = c_oracle.logp_grad(data, det, 3.77)
        assert used >= 1
        assert_close(c_lp, lp.detach().numpy(), rtol=1e-12, what="C oracle logp", axis=None)
        names = [k for k in c_oracle.INPUTS if lor or k != "fwhm_L"]
        got = np.concatenate([c_g[k] for k in names], axis=1)
        ref = np.concatenate([t[k].grad.numpy() for k in names], axis=1)
        assert_close(got, ref, rtol=1e-10, what="C oracle gradient")


def test_c_oracle_single_dataset():
    det, data = _case(True, 3)
    spec = mr.make_spec("emission", 3)
    t = {k: torch.tensor(v, requires_grad=True) for k, v in det.items()}
    lp = mr.logp_from_inputs(spec, {"emission": data["emission"]}, t)
    lp.sum().backward()
    c_lp, c_g, _ = c_oracle.logp_grad({"emission": data["emission"]}, det, 3.77, n_threads=2)
    assert_close(c_lp, lp.detach().numpy(), rtol=1e-12, what="logp", axis=None)
    assert_close(c_g["tspin"], t["tspin"].grad.numpy(), what="d/dtspin")
