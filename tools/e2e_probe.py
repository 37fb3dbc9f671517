"""Times the host entry points with pinned buffers on the bench workload: slot blocks, packed
rows, with and without baseline coefficients, several pipeline schedules; prints the box's raw
pinned copy bandwidth next to it (development aid).  usage: python tools/e2e_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caribou_hi_b200 import Engine, PinnedArray, synthetic

B, N, S = 100000, 4, 1000
kind = "emission_absorption_physical"


def build(with_coeff):
    eng = Engine(kind, N, prior_amplitude=1e21)
    v = np.linspace(-100, 100, S); rng = np.random.default_rng(0)
    kw = dict(basis=np.full((1, S), 0.1)) if with_coeff else {}
    eng.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1, **kw),
                    absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01, **kw))
    th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal"]
    return eng, eng.pack(synthetic.draw_free(kind, N, B, rng, thermal_fn=th))


def run(eng, fn, reps=20):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


keep = []
for with_coeff in (True, False):
    eng, theta = build(with_coeff)
    rows = eng.row_slots
    pth, prow = PinnedArray((B, 10, N)), PinnedArray((B, rows.size, N))
    pth.array[...] = theta; prow.array[...] = theta[:, rows, :]
    pins = [PinnedArray((B, 1)) for _ in range(4)]
    pins[0].array[...] = 0; pins[1].array[...] = 0
    plp, pgr, pgrr = PinnedArray((B,)), PinnedArray((B, 10, N)), PinnedArray((B, rows.size, N))
    keep += [pth, prow, plp, pgr, pgrr] + pins
    ce, ca = (pins[0].array, pins[1].array) if with_coeff else (None, None)
    out = {"logp": plp.array, "grad": pgr.array, "gcoeff_e": pins[2].array, "gcoeff_a": pins[3].array}
    outr = dict(out, grad=pgrr.array)
    for env in (None, "3", "6"):
        if env is None:
            os.environ.pop("CHI_PIPE_PIECES", None)
        else:
            os.environ["CHI_PIPE_PIECES"] = env
        a = run(eng, lambda: eng.logp_grad(pth.array, ce, ca, out=out))
        b = run(eng, lambda: eng.logp_grad_rows(prow.array, ce, ca, out=outr))
        print(f"coeff={with_coeff} pieces={env or 'ramp'}: slots {a:.3f} ms, rows {b:.3f} ms, kernels {eng.stage_ms(9).sum():.3f} ms", flush=True)
    eng.close()

x = torch.empty(B * 10 * N, dtype=torch.float64, device="cuda")
hp = torch.empty(B * 10 * N, dtype=torch.float64).pin_memory()
for name, fn in (("H2D pinned", lambda: x.copy_(hp, non_blocking=True)), ("D2H pinned", lambda: hp.copy_(x, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {x.numel()*8/dt/1e9:.1f} GB/s ({dt*1e3:.3f} ms for {x.numel()*8/1e6:.1f} MB)")
