import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caribou_hi_b200 import EmissionAbsorptionPhysicalModel, SpecData
v = np.linspace(-60, 60, 512)
for tick in (2, 0):
    m = EmissionAbsorptionPhysicalModel({"emission": SpecData(v, np.full(512, 5.0), 0.5), "absorption": SpecData(v, np.full(512, 0.2), 0.05)}, 9, baseline_degree=0)
    m.add_priors(prior_fwhm_L=1.0); m.add_likelihood()
    m.engine.set_option("nuts_tick", tick)
    n0 = m.engine.launch_count()
    post, st = m.sample(draws=10, tune=30, chains=4, seed=4)
    print("N=9 nuts_tick", tick, "launches", m.engine.launch_count() - n0, "accept", float(st["accept"].mean()), "depth", float(st["depth"].mean()))
    m.engine.close()
