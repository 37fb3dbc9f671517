"""Tiny runs of every form of the round's new kernels -- the cluster form of the one-launch evaluation, the
two-launch forms of the large-batch kernels, the sampler's chain kernels -- as a target for compute-sanitizer
where it is available (it is closed on the pool this was developed on) or for a quick look after a change.
usage: [compute-sanitizer --tool memcheck] python tools/sanitize_target.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caribou_hi_b200 import EmissionAbsorptionModel, EmissionAbsorptionPhysicalModel, Engine, SpecData, synthetic

rng = np.random.default_rng(0)
# one-launch evaluation, blocks of a draw as a cluster (and the global-memory form)
kind, N, S, B = "emission_absorption_physical", 4, 512, 6
eng = Engine(kind, N, lorentzian=True, prior_fwhm_L=1.0, prior_amplitude=1e21)
v = np.linspace(-100, 100, S)
eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1, basis=np.full((1, S), 0.1)),
                absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01, basis=np.full((1, S), 0.01)))
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
theta = eng.pack(synthetic.draw_free(kind, N, B, rng, lorentzian=True, thermal_fn=th))
eng.set_option("small_path", 2)
for cl in (0, 1):
    eng.set_option("small_cluster", cl)
    lp, g, _, _ = eng.logp_grad(theta, np.zeros((B, 1)), np.zeros((B, 1)))
print("one-launch ok", np.isfinite(lp).sum())
eng.set_option("small_path", 0)
# large batch: EA_TAME + EA_FIXUP (shared axis, Gaussian) with a few marked draws
eng2 = Engine("physics", 3, lorentzian=False)
eng2.set_spectra(emission=dict(spectral=v, brightness=np.ones(S), noise=0.1), absorption=dict(spectral=v, brightness=np.full(S, 0.2), noise=0.01))
Bb = 6000
inp = {"velocity": rng.normal(0, 8, (Bb, 3)), "fwhm2": rng.uniform(5, 200, (Bb, 3)), "fwhm_L": np.zeros((Bb, 3)),
       "tau_total": rng.uniform(0.05, 8, (Bb, 3)), "tspin": rng.uniform(30, 3000, (Bb, 3)),
       "filling_factor": rng.uniform(0.1, 1.0, (Bb, 3)), "absorption_weight": rng.uniform(0.3, 1.2, (Bb, 3))}
inp["tau_total"][::50, 0] = -0.3
lp, g, _, _ = eng2.logp_grad(eng2.pack(inp))
print("two-launch ok", np.isfinite(lp).sum())
eng2.close(); eng.close()
# sampler: every chain-kernel form, a handful of iterations
ve, va = np.linspace(-40, 40, 96), np.linspace(-20, 20, 48)
for tick in (3, 4, 5, 6):
    m = EmissionAbsorptionModel({"emission": SpecData(ve, np.full(96, 5.0), 0.5), "absorption": SpecData(va, np.full(48, 0.2), 0.05)},
                                2, baseline_degree=1)
    m.add_priors(); m.add_likelihood()
    m.engine.set_option("nuts_tick", tick)
    post, st = m.sample(draws=10, tune=40, chains=3, seed=4)
    print("nuts_tick", tick, "ok", float(st["accept"].mean()))
    m.engine.close()
