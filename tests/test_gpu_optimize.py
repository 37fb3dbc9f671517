"""Optimize: the n_clouds sweep of bayes_spec.Optimize (optimization.ipynb cells 8-13) on the device sampler."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sweep_recovers_the_number_of_clouds(cuda_device):
    """Data simulated from two well separated clouds: BIC falls steeply from the null hypothesis to n = 1 and to
    n = 2, adding a third and a fourth cloud does not pay, the sweep stops early and picks two."""
    import torch

    from caribou_hi_b200 import AbsorptionModel, Optimize, SpecData

    rng = np.random.default_rng(3)
    S = 200
    v = np.linspace(-40.0, 40.0, S)
    sim = AbsorptionModel({"absorption": SpecData(v, np.zeros(S), 0.01)}, 2, baseline_degree=0)
    sim.add_priors(prior_tau_total=10.0)
    sim.add_likelihood()
    truth = {"fwhm2_norm": np.array([[0.05, 0.08]]), "velocity_norm": np.array([[-1.2, 1.0]]),
             "log10_nHI_norm": np.zeros((1, 2)), "log10_n_alpha_norm": np.zeros((1, 2)),
             "tau_total_norm": np.array([[0.6, 0.4]]), "tkin_factor": np.full((1, 2), 0.5),
             "baseline_absorption_norm": np.zeros((1, 1))}
    mu = sim.predict(truth)["absorption"][0]
    assert np.all(np.isfinite(mu)) and mu.max() > 0.1
    data = {"absorption": SpecData(v, mu + 0.01 * rng.standard_normal(S), 0.01)}
    devices = tuple(range(min(2, torch.cuda.device_count())))
    opt = Optimize(AbsorptionModel, data, max_n_clouds=5, baseline_degree=0, seed=5, verbose=True, devices=devices)
    opt.add_priors(prior_tau_total=10.0)
    opt.add_likelihood()
    assert sorted(opt.models) == [1, 2, 3, 4, 5] and opt.models[4].n_clouds == 4     # notebook cell 10
    assert Optimize.n_parameters(opt.models[2]) == 2 * 6 + 1
    best = opt.optimize(bic_threshold=10.0, sample_kwargs={"draws": 150, "tune": 250, "chains": 8, "seed": 11,
                                                           "cores": 8, "nuts_kwargs": {"target_accept": 0.8}})
    assert best is opt.best_model and best.n_clouds == 2, opt.bics
    assert opt.bics[0] > opt.bics[1] > opt.bics[2]
    assert 5 not in opt.bics                  # stopped early: two counts in a row did not improve the BIC
    assert abs(opt.bics[0] - opt.null_bic()) < 1e-9
    post = opt.results[2]["posterior"]
    assert post["velocity_norm"].shape == (8, 150, 2)
    # the two recovered line centres
    vel = np.sort(np.median(post["velocity_norm"].reshape(-1, 2), axis=0) * 10.0)
    assert np.all(np.abs(vel - np.array([-12.0, 10.0])) < 1.0), vel
