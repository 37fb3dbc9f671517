"""Host-call vs device-call latency of logp_grad at several batch sizes (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine, PinnedArray, synthetic

N, S = 4, 1000
kind = "emission_absorption_physical"
eng = Engine(kind, N, prior_amplitude=1e21)
v = np.linspace(-100, 100, S); rng = np.random.default_rng(0)
eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1),
                absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01))
th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
Bmax = 100000
theta = eng.pack(synthetic.draw_free(kind, N, Bmax, rng, thermal_fn=th))
pth, plp, pgr = PinnedArray((Bmax, 10, N)), PinnedArray((Bmax,)), PinnedArray((Bmax, 10, N))
pth.array[...] = theta
th_d = torch.from_numpy(theta).cuda(); lp_d = torch.empty(Bmax, dtype=torch.float64, device="cuda"); gr_d = torch.empty_like(th_d)
for B in (1024, 4096, 16384, 32768, 100000):
    os.environ["CHI_PIPE_PIECES"] = "1"
    out = {"logp": plp.array[:B], "grad": pgr.array[:B]}
    def host(): eng.logp_grad(pth.array[:B], out=out)
    def dev():
        eng.logp_grad_dev(B, th_d, lp_d, gr_d); eng.synchronize()
    res = []
    for fn in (host, dev):
        for _ in range(3): fn()
        t0 = time.perf_counter()
        for _ in range(20): fn()
        res.append((time.perf_counter() - t0) / 20 * 1e3)
    mb = B * 10 * N * 8 / 1e6
    print(f"B={B}: host call {res[0]:.3f} ms, device call + sync {res[1]:.3f} ms, difference {res[0]-res[1]:.3f} ms for {mb:.1f} MB each way"
          f" -> {2*mb/ (res[0]-res[1]) :.1f} GB/s effective", flush=True)
