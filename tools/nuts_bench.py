"""Times the chain-batched NUTS on BASELINE configs[0] and configs[2] shapes (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caribou_hi_b200 import EmissionAbsorptionModel, EmissionAbsorptionPhysicalModel, SpecData

def run(cls, N, S_e, S_a, chains, tune, draws, **priors):
    rng = np.random.default_rng(1)
    ve, va = np.linspace(-60, 60, S_e), np.linspace(-30, 30, S_a)
    sim = cls({"emission": SpecData(ve, np.zeros(S_e), 0.1), "absorption": SpecData(va, np.zeros(S_a), 0.01)}, N, baseline_degree=0)
    sim.add_priors(**priors); sim.add_likelihood()
    truth = sim.sample_prior(1, np.random.default_rng(7))
    for _ in range(20):
        mu = sim.predict(truth)
        if all(np.all(np.isfinite(m)) for m in mu.values()): break
        truth = sim.sample_prior(1, rng)
    data = {"emission": SpecData(ve, mu["emission"][0] + 0.1 * rng.standard_normal(S_e), 0.1),
            "absorption": SpecData(va, mu["absorption"][0] + 0.01 * rng.standard_normal(S_a), 0.01)}
    model = cls(data, N, baseline_degree=0)
    model.add_priors(**priors); model.add_likelihood()
    if os.environ.get("NUTS_TICK"):
        model.engine.set_option("nuts_tick", int(os.environ["NUTS_TICK"]))
    l0 = model.engine.launch_count()
    t0 = time.perf_counter()
    post, stats = model.sample(draws=draws, tune=tune, chains=chains, seed=2, init_point=truth, jitter=0.05,
                               pool_mass=bool(os.environ.get("POOL")), lock_step=bool(os.environ.get("LOCK_STEP")))
    dt = time.perf_counter() - t0
    nl = (2.0 ** stats["depth"]).sum()  # upper bound on leapfrogs of the kept draws
    print(f"{cls.__name__} N={N} S={S_e}+{S_a} chains={chains}: {tune}+{draws} iterations in {dt:.2f} s "
          f"({(tune+draws)/dt:.0f} it/s), mean depth {stats['depth'].mean():.2f}, accept {stats['accept'].mean():.2f}, "
          f"divergent {stats['diverging'].mean():.3f}, launches {model.engine.launch_count()-l0}", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:  # many chains in lock step: python tools/nuts_bench.py <chains> [tune draws]
        c = int(sys.argv[1]); t = int(sys.argv[2]) if len(sys.argv) > 2 else 200; d = int(sys.argv[3]) if len(sys.argv) > 3 else 200
        run(EmissionAbsorptionModel, 2, 334, 166, c, t, d)
        sys.exit(0)
    run(EmissionAbsorptionModel, 2, 334, 166, 8, 500, 500)
    run(EmissionAbsorptionPhysicalModel, 8, 2048, 2048, 64, 200, 200, prior_fwhm_L=1.0)
