"""Sampler efficiency diagnostics: tree depth / acceptance / step size on a flat-likelihood
(prior-only) model and on an informative one."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caribou_hi_b200 import EmissionModel, AbsorptionModel, SpecData

rng = np.random.default_rng(0)
v = np.linspace(-30, 30, 64)
model = EmissionModel({"emission": SpecData(v, rng.normal(0, 1, 64), 1.0e7)}, 2, baseline_degree=0)
model.add_priors(); model.add_likelihood()
for ta in (0.8, 0.9):
    post, stats = model.sample(draws=300, tune=500, chains=64, seed=11, target_accept=ta)
    d = stats["depth"]
    print(f"prior-only target={ta}: depth mean {d.mean():.2f} hist {np.bincount(d.astype(int).ravel(), minlength=11)[:11]}"
          f" accept {stats['accept'].mean():.3f} keys {sorted(stats)}", flush=True)
    for k in ("step_size", "stepsize", "eps"):
        if k in stats:
            print("  step size", np.percentile(stats[k], [5, 50, 95]))
