"""Prints the per-piece stage timeline of one pipelined host call (CHI_TIMELINE=1)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caribou_hi_b200 import Engine, PinnedArray, synthetic
B, N, S = 100000, 4, 1000
kind = "emission_absorption_physical"
eng = Engine(kind, N, prior_amplitude=1e21)
v = np.linspace(-100, 100, S); rng = np.random.default_rng(0)
eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1),
                absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01))
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
theta = eng.pack(synthetic.draw_free(kind, N, B, rng, thermal_fn=th))
pth, plp, pgr = PinnedArray((B, 10, N)), PinnedArray((B,)), PinnedArray((B, 10, N))
pth.array[...] = theta
out = {"logp": plp.array, "grad": pgr.array}
for _ in range(3): eng.logp_grad(pth.array, out=out)
os.environ["CHI_TIMELINE"] = "1"
t0 = time.perf_counter(); eng.logp_grad(pth.array, out=out); print("wall %.3f ms" % ((time.perf_counter() - t0) * 1e3))
