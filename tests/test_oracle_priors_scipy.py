"""The prior log-densities and transform Jacobians restated in oracle/priors_ref.py (PyMC is not
installable here) against an independent implementation: the distribution family and
parameters are the ones THE REFERENCE'S SOURCE declares (tests/golden/ref_free_rvs.json, made by
make_golden_ref.py), the log-density is scipy.stats', the Jacobian a numerical derivative."""
import json
import os

import numpy as np
import pytest
from scipy import stats

from oracle import models_ref as mr
from oracle import priors_ref as pr

HERE = os.path.dirname(os.path.abspath(__file__))


def _scipy_logpdf(info, x):
    d, p = info["dist"], info["params"]
    if d == "Normal":
        return stats.norm.logpdf(x, loc=np.asarray(p["mu"]), scale=np.asarray(p["sigma"]))
    if d == "HalfNormal":
        return stats.halfnorm.logpdf(x, scale=p["sigma"])
    if d == "ChiSquared":
        return stats.chi2.logpdf(x, df=p["nu"])
    if d == "Beta":
        return stats.beta.logpdf(x, p["alpha"], p["beta"])
    if d == "Uniform":
        return stats.uniform.logpdf(x, loc=p["lower"], scale=p["upper"] - p["lower"])
    if d == "TruncatedNormal":
        a = (p["lower"] - p["mu"]) / p["sigma"]
        return stats.truncnorm.logpdf(x, a, np.inf, loc=p["mu"], scale=p["sigma"])
    raise KeyError(d)


def _samples(info, rng, n):
    d = info["dist"]
    if d == "Normal":
        return rng.normal(0.0, 2.0, n)
    if d in ("HalfNormal", "ChiSquared"):
        return 10 ** rng.uniform(-3, 1.2, n)
    if d in ("Beta", "Uniform"):
        return rng.uniform(1e-4, 1 - 1e-4, n)
    return rng.uniform(0.01, 4.0, n)  # TruncatedNormal(lower=0)


def test_densities_and_jacobians_against_scipy():
    with open(os.path.join(HERE, "golden", "ref_free_rvs.json")) as f:
        meta = json.load(f)
    rng = np.random.default_rng(5)
    checked = set()
    for case, rvs in meta.items():
        kind = case.rsplit("_", 1)[0]
        spec = mr.make_spec(kind, 3, **({"prior_sigma_log10_NHI": 0.5} if kind == "emission_absorption" else {}))
        for name, info in rvs.items():
            if name.startswith("baseline_") or name.startswith("log10_wt_ff_"):
                continue  # baseline: Normal(0, sigma) per coefficient; the coupled Normal is below
            x = _samples(info, rng, 40)
            got = pr.density(spec, name, x, {})
            np.testing.assert_allclose(got, _scipy_logpdf(info, x), rtol=1e-12, atol=1e-12, err_msg=f"{case} {name}")
            # transform: x = backward(u); log|dx/du| by central differences
            u = pr.forward(name, x)
            np.testing.assert_allclose(pr.backward(name, u), x, rtol=1e-12)
            h = 1e-6
            num = np.log(np.abs(pr.backward(name, u + h) - pr.backward(name, u - h)) / (2 * h))
            np.testing.assert_allclose(pr.log_jacobian(name, u, x), num, rtol=0, atol=2e-6, err_msg=f"{case} {name}")
            checked.add(info["dist"])
    assert checked == {"Normal", "HalfNormal", "ChiSquared", "Beta", "Uniform", "TruncatedNormal"}


@pytest.mark.parametrize("kind", ["emission_absorption", "emission_absorption_physical"])
def test_coupled_weight_prior_against_scipy(kind):
    """log10_wt_ff_tkin ~ Normal(mu = -ln10 sigma^2/2 - log10(tkin or fwhm2_thermal), sigma)
    (emission_absorption_model.py:45-58, emission_absorption_physical_model.py:41-54): the mean
    recorded from the reference's source for draw 0 of the fixture."""
    key = "tkin" if kind == "emission_absorption" else "fwhm2_thermal"
    name = f"log10_wt_ff_{key}"
    with open(os.path.join(HERE, "golden", "ref_free_rvs.json")) as f:
        info = json.load(f)[f"{kind}_gauss"][name]
    z = np.load(os.path.join(HERE, "golden", f"ref_{kind}_gauss.npz"))
    spec = mr.make_spec(kind, 3, prior_sigma_log10_NHI=0.5)
    x = z[f"free_{name}"][0]
    det = {key: z[f"det_{key}"][0]}
    got = pr.density(spec, name, x, det)
    np.testing.assert_allclose(got, _scipy_logpdf(info, x), rtol=1e-12, atol=1e-12)
