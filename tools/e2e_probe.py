"""Times Engine.logp_grad with pinned buffers for several pipeline piece counts, and the raw
pinned H2D/D2H bandwidth (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine, PinnedArray, synthetic

B, N, S = 100000, 4, 1000
kind = "emission_absorption_physical"
pieces = os.environ.get("CHI_PIPE_PIECES", "default")
eng = Engine(kind, N, prior_amplitude=1e21)
v = np.linspace(-100, 100, S); rng = np.random.default_rng(0)
eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1, basis=np.full((1, S), 0.1)),
                absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01, basis=np.full((1, S), 0.01)))
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
theta = PinnedArray((B, 10, N)); theta.array[...] = eng.pack(synthetic.draw_free(kind, N, B, rng, thermal_fn=th))
pins = [PinnedArray((B, 1)) for _ in range(4)]
pins[0].array[...] = 0; pins[1].array[...] = 0
plp, pgr = PinnedArray((B,)), PinnedArray((B, 10, N))  # keep the owners alive
out = {"logp": plp.array, "grad": pgr.array, "gcoeff_e": pins[2].array, "gcoeff_a": pins[3].array}
for _ in range(3):
    eng.logp_grad(theta.array, pins[0].array, pins[1].array, out=out)
t0 = time.perf_counter()
for _ in range(10):
    eng.logp_grad(theta.array, pins[0].array, pins[1].array, out=out)
dt = (time.perf_counter() - t0) / 10
print(f"pieces={pieces}: e2e {dt*1e3:.3f} ms", flush=True)
if pieces == "default":
    x = torch.empty(B * 10 * N, dtype=torch.float64, device="cuda")
    h = torch.from_numpy(theta.array.reshape(-1))
    hp = torch.empty(B * 10 * N, dtype=torch.float64).pin_memory()
    for name, fn in (("H2D pinned", lambda: x.copy_(hp, non_blocking=True)), ("D2H pinned", lambda: hp.copy_(x, non_blocking=True))):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f"{name}: {x.numel()*8/dt/1e9:.1f} GB/s ({dt*1e3:.3f} ms for {x.numel()*8/1e6:.1f} MB)")
