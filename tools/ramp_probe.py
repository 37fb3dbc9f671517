"""e2e time of the packed-row host call for several piece schedules (CHI_PIPE_RAMP)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caribou_hi_b200 import Engine, PinnedArray, synthetic
B, N, S = 100000, 4, 1000
kind = "emission_absorption_physical"
eng = Engine(kind, N, prior_amplitude=1e21)
v = np.linspace(-100, 100, S); rng = np.random.default_rng(0)
eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1, basis=np.full((1, S), 0.1)),
                absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01, basis=np.full((1, S), 0.01)))
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
rows = eng.pack_rows(synthetic.draw_free(kind, N, B, rng, thermal_fn=th))
prow, plp, pgr = PinnedArray(rows.shape), PinnedArray((B,)), PinnedArray(rows.shape)
pins = [PinnedArray((B, 1)) for _ in range(4)]
prow.array[...] = rows; pins[0].array[...] = 0; pins[1].array[...] = 0
out = {"logp": plp.array, "grad": pgr.array, "gcoeff_e": pins[2].array, "gcoeff_a": pins[3].array}
prow2, plp2, pgr2 = PinnedArray(rows.shape), PinnedArray((B,)), PinnedArray(rows.shape)
pins2 = [PinnedArray((B, 1)) for _ in range(2)]
prow2.array[...] = rows
outs = [out, {"logp": plp2.array, "grad": pgr2.array, "gcoeff_e": pins2[0].array, "gcoeff_a": pins2[1].array}]
ins = [prow.array, prow2.array]


def run_async(n):
    prev = None
    for i in range(n):
        tk = eng.logp_grad_rows_async(ins[i % 2], pins[0].array, pins[1].array, outs[i % 2])
        if prev is not None:
            eng.wait(prev)
        prev = tk
    eng.wait(prev)


for ramp in sys.argv[1:]:
    if os.environ.get("ASYNC"):
        os.environ["CHI_PIPE_RAMP"] = ramp
        best = 1e9
        for rep in range(3):
            run_async(4)
            t0 = time.perf_counter()
            run_async(30)
            best = min(best, (time.perf_counter() - t0) / 30 * 1e3)
        print(f"async ramp {ramp}: {best:.3f} ms", flush=True)
        continue
    os.environ["CHI_PIPE_RAMP"] = ramp
    best = 1e9
    for rep in range(3):
        for _ in range(3): eng.logp_grad_rows(prow.array, pins[0].array, pins[1].array, out=out)
        t0 = time.perf_counter()
        for _ in range(20): eng.logp_grad_rows(prow.array, pins[0].array, pins[1].array, out=out)
        best = min(best, (time.perf_counter() - t0) / 20 * 1e3)
    print(f"ramp {ramp}: {best:.3f} ms", flush=True)
