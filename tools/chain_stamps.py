"""Phase times of the per-chain NUTS kernel (development aid): cycles block 0 spends in the chain's state
machine and in the evaluation, per tick, plus the evaluation's own phase stamps of the last tick.
usage: python tools/chain_stamps.py [chains]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import EmissionAbsorptionModel, SpecData

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rng = np.random.default_rng(1)
S_e, S_a, N = 334, 166, 2
ve, va = np.linspace(-60, 60, S_e), np.linspace(-30, 30, S_a)
sim = EmissionAbsorptionModel({"emission": SpecData(ve, np.zeros(S_e), 0.1), "absorption": SpecData(va, np.zeros(S_a), 0.01)}, N, baseline_degree=0)
sim.add_priors(); sim.add_likelihood()
truth = sim.sample_prior(1, np.random.default_rng(7))
mu = sim.predict(truth)
data = {"emission": SpecData(ve, mu["emission"][0] + 0.1 * rng.standard_normal(S_e), 0.1),
        "absorption": SpecData(va, mu["absorption"][0] + 0.01 * rng.standard_normal(S_a), 0.01)}
model = EmissionAbsorptionModel(data, N, baseline_degree=0)
model.add_priors(); model.add_likelihood()
stamps = torch.zeros(16, dtype=torch.int64, device="cuda")
model.engine.set_option("nuts_tick", int(os.environ.get("NUTS_TICK", 3)))
model.engine.set_option("small_stamps", stamps.data_ptr())
import time
t0 = time.perf_counter()
post, stats = model.sample(draws=100, tune=100, chains=chains, seed=2, init_point=truth, jitter=0.05)
dt = time.perf_counter() - t0
st = stamps.cpu().numpy()
n = max(int(st[14]), 1)
mhz = 1965.0
print(f"{chains} chains: 200 iterations in {dt:.2f} s; block 0: {n} ticks, state machine {st[12]/n/mhz:.2f} us/tick, "
      f"evaluation {st[13]/n/mhz:.2f} us/tick")
print("evaluation phases of the last tick (us from its start): map %.2f jac %.2f [emission loop %.2f, its sums %.2f] chan %.2f joined %.2f | logp %.2f grad %.2f"
      % (st[0]/mhz, st[9]/mhz, st[7]/mhz, st[15]/mhz, st[2]/mhz, st[3]/mhz, st[5]/mhz, st[6]/mhz))
