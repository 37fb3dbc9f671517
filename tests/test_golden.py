"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py):
  * CPU: the oracle still reproduces them (guards the checker against drift);
  * GPU: the CUDA path reproduces them through the C ABI."""

import glob
import os

import numpy as np
import pytest

from helpers import assert_close, engine_for
from oracle import models_ref as mr

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    kind = str(z["kind"])
    lor = bool(z["lorentzian"])
    kw = {"prior_fwhm_L": 1.0} if lor else {}
    if kind == "emission_absorption":
        kw["prior_sigma_log10_NHI"] = 0.5
    spec = mr.make_spec(kind, 3, **kw)
    data = {}
    for key in ("emission", "absorption"):
        if f"data_{key}_spectral" in z:
            data[key] = mr.OracleData(z[f"data_{key}_spectral"], z[f"data_{key}_brightness"], z[f"data_{key}_noise"])
    free = {k[5:]: z[k] for k in z.files if k.startswith("free_") and not k.startswith("free_baseline_")}
    base = {k[5:]: z[k] for k in z.files if k.startswith("free_baseline_")}
    return z, spec, data, free, base


def test_fixtures_present():
    assert len(FILES) == 12


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    from helpers import oracle_logp_grad

    z, spec, data, free, base = _load(path)
    lp, grads = oracle_logp_grad(spec, data, free, base)
    assert_close(lp, z["logp"], rtol=1e-13, what="logp", axis=None)
    fin = np.isfinite(z["logp"])
    for k, g in grads.items():
        assert_close(g[fin], z[f"grad_{k}"][fin], rtol=1e-12, what=k, axis=None)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_reproduces_golden(cuda_device, path):
    z, spec, data, free, base = _load(path)
    eng = engine_for(spec, data, baseline_degree=1)
    theta = eng.pack(free)
    lp, grad, gce, gca = eng.logp_grad(theta, base.get("baseline_emission_norm"), base.get("baseline_absorption_norm"))
    assert_close(lp, z["logp"], what="logp", axis=None)
    fin = np.isfinite(z["logp"])
    got = eng.unpack_grad(grad)
    names = sorted(got)
    x = [got[n][fin] for n in names]
    r = [z[f"grad_{n}"][fin] for n in names]
    if gce is not None:
        x.append(gce[fin]); r.append(z["grad_baseline_emission_norm"][fin])
    if gca is not None:
        x.append(gca[fin]); r.append(z["grad_baseline_absorption_norm"][fin])
    assert_close(np.concatenate(x, axis=1), np.concatenate(r, axis=1), what="gradient")
    mu_e, mu_a = eng.spectra(theta, base.get("baseline_emission_norm"), base.get("baseline_absorption_norm"))
    if mu_e is not None:
        assert_close(mu_e, z["mu_emission"], what="mu_e")
    if mu_a is not None:
        assert_close(mu_a, z["mu_absorption"], what="mu_a")
    det = eng.deterministics(theta)
    for k in z.files:
        if k.startswith("det_") and k[4:] in det:
            assert_close(det[k[4:]], z[k], rtol=1e-12, what=k)
    eng.close()
