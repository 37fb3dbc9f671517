"""Where does a gradient component lose its digits?  For one draw of tests/test_gpu_parity.py::
test_branch_edges_of_the_parameter_maps: the CUDA gradient with respect to the free RVs and with respect to
the seven likelihood inputs (CHI_MODEL_PHYSICS engine at the draw's deterministic values), fp64 autograd and
the 60-digit derivative of both.  usage: python tools/grad_probe.py [kind draw cloud]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpmath as mp
import torch
from helpers import engine_for, make_case, mp_derivative, oracle_logp_grad
from oracle import models_ref as mr
from oracle.backend import MpmathBackend as MB
from caribou_hi_b200 import Engine

kind = sys.argv[1] if len(sys.argv) > 1 else "emission"
b0 = int(sys.argv[2]) if len(sys.argv) > 2 else 591
n0 = int(sys.argv[3]) if len(sys.argv) > 3 else 2
B, N = 768, 3
spec, data, free, baseline = make_case(kind, N, 96, 64, B, seed=909, lorentzian=True)
rng = np.random.default_rng(910)
amp = [k for k in free if k in ("TB_fwhm_norm", "ff_NHI_norm", "tau_total_norm", "NHI_fwhm2_thermal_norm")][0]
free[amp] = free[amp] * 10 ** rng.uniform(-3, 3, size=(B, 1))
free["fwhm2_norm"] = free["fwhm2_norm"] * 10 ** rng.uniform(-2, 1, size=(B, 1))
for k in ("filling_factor_norm", "tkin_factor_norm", "fwhm2_thermal_fraction_norm"):
    if k in free:
        free[k][: B // 8] = 0.0
        free[k][B // 8 : B // 4] = 1.0
free["fwhm_L_norm"][B // 2 :: 7] = 0.0
one = lambda d: {k: np.array(v[b0 : b0 + 1]) for k, v in d.items()}
f1, bl1 = one(free), one(baseline)
lp_ref, g_ref = oracle_logp_grad(spec, data, f1, bl1)
eng = engine_for(spec, data, 0)
lp, grad, _, _ = eng.logp_grad(eng.pack(f1), bl1.get("baseline_emission_norm"), bl1.get("baseline_absorption_norm"))
got = eng.unpack_grad(grad)
print(f"free RVs of draw {b0}, cloud {n0} ({kind}): name, 60-digit, cuda rel err, autograd rel err")
for name in sorted(got):
    n = n0 if got[name].shape[1] > 1 else 0
    ex = float(mp_derivative(spec, data, f1, bl1, 0, name, n))
    print(f"  {name:24s} {ex: .12e}  cuda {abs(got[name][0, n] - ex) / abs(ex):.2e}  autograd {abs(g_ref[name][0, n] - ex) / abs(ex):.2e}")
eng.close()

# the seven likelihood inputs at the draw's deterministic values
det = {k: np.broadcast_to(v, (1, N)).copy() for k, v in mr.cloud_inputs(spec, f1).items() if np.ndim(v) > 0}
names = ["velocity", "fwhm2", "fwhm_L", "tau_total", "tspin", "filling_factor", "absorption_weight"]
inp = {k: det[k] if k in det else np.ones((1, N)) for k in names}
pe = Engine("physics", N, lorentzian=True, bg_temp=spec["bg_temp"], has_emission=mr.has_emission(spec), has_absorption=mr.has_absorption(spec))
sets = {}
for key, d in data.items():
    basis = np.stack([d.spectral_norm**i / (i + 1.0) ** i * d._brightness_scale for i in range(1)])
    sets[key] = dict(spectral=d.spectral, brightness=d.brightness, noise=d.noise, basis=basis, offset=d._brightness_offset)
pe.set_spectra(emission=sets.get("emission"), absorption=sets.get("absorption"))
lp2, g2, _, _ = pe.logp_grad(pe.pack(inp), bl1.get("baseline_emission_norm"), bl1.get("baseline_absorption_norm"))
gp = pe.unpack_grad(g2)
it = {k: torch.tensor(v, requires_grad=True) for k, v in inp.items()}
bt = {k: torch.tensor(v) for k, v in bl1.items()}
mr.logp_from_inputs(spec, data, it, bt).sum().backward()
mp.mp.dps = 60
def f_in(name, delta):
    d = {k: MB.asarray(v) for k, v in inp.items()}
    d[name] = d[name].copy(); d[name][0, n0] = d[name][0, n0] + delta
    return mr.logp_from_inputs(spec, data, d, {k: MB.asarray(v) for k, v in bl1.items()})[0]
h = mp.mpf(10) ** (-25)
print(f"likelihood inputs: logp cuda {lp2[0]!r} vs model-level {lp[0]!r}")
for name in names:
    if it[name].grad is None:
        continue
    ex = float((f_in(name, h) - f_in(name, -h)) / (2 * h))
    if ex == 0.0:
        continue
    print(f"  {name:24s} {ex: .12e}  cuda {abs(gp[name][0, n0] - ex) / abs(ex):.2e}  autograd {abs(float(it[name].grad[0, n0]) - ex) / abs(ex):.2e}")
pe.close()

# forward values of the map: CUDA deterministics and the fp64 oracle against 60 digits
eng = engine_for(spec, data, 0)
dd = eng.deterministics(eng.pack(f1))
d64 = mr.cloud_inputs(spec, f1)
dmp = mr.cloud_inputs(spec, {k: MB.asarray(v) for k, v in f1.items()})
print("map values: name, 60-digit, cuda rel err, fp64 oracle rel err")
for k in ("tkin", "tspin", "filling_factor", "tau_total"):
    if k in dd and k in dmp:
        ex = dmp[k][0, n0] if np.ndim(dmp[k]) == 2 else dmp[k][0]
        o64 = np.broadcast_to(d64[k], (1, N))[0, n0]
        print(f"  {k:16s} {float(ex): .15e}  cuda {abs(float((mp.mpf(float(dd[k][0, n0])) - ex) / ex)):.2e}  oracle {abs(float((mp.mpf(float(o64)) - ex) / ex)):.2e}")
eng.close()

# chain by hand for filling_factor_norm of the emission models: which side loses the digits?
if kind in ("emission", "emission_absorption"):
    fwhm2 = det["fwhm2"][0, n0]; TB = det["TB_fwhm"][0, n0]; tspin = det["tspin"][0, n0]; ff = det["filling_factor"][0, n0]
    tkin_min = TB / np.sqrt(fwhm2); ff_min = tkin_min / tspin; x = ff_min / ff
    G = np.sqrt(2 * np.pi) / (2 * np.sqrt(2 * np.log(2)))
    dtau = np.sqrt(fwhm2) * G * (1 / (1 - x)) * (-x / ff) * (1 - ff_min)
    ex = float(mp_derivative(spec, data, f1, bl1, 0, "filling_factor_norm", n0))
    for who, fb, tb in (("cuda kernel adjoints", gp["filling_factor"][0, n0], gp["tau_total"][0, n0]),
                        ("autograd adjoints", float(it["filling_factor"].grad[0, n0]), float(it["tau_total"].grad[0, n0]))):
        hand = fb * (1 - ff_min) + tb * dtau
        print(f"  chain by hand with {who}: {hand!r} rel err {abs(hand - ex) / abs(ex):.2e} (terms {fb * (1 - ff_min)!r} {tb * dtau!r})")
