"""Generates tests/golden/*.npz: seeded inputs and the CPU oracle's outputs (logp, gradients
by torch autograd, predicted spectra, deterministics) for every model class x {Gaussian,
pseudo-Voigt}.  These vectors guard the ORACLE against drift and give the GPU tests further
fixed cases; the vectors produced by the reference's own source are the `ref_*` files made by
make_golden_ref.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_case, oracle_logp_grad  # noqa: E402
from oracle import models_ref as mr  # noqa: E402

CASES = [(kind, lor) for kind in mr.KINDS for lor in (False, True)]


def main():
    for i, (kind, lor) in enumerate(CASES):
        spec, data, free, baseline = make_case(kind, 3, 64, 48, 6, seed=500 + i, lorentzian=lor, baseline_degree=1)
        lp, grads = oracle_logp_grad(spec, data, free, baseline)
        mu = mr.predict(spec, data, free, baseline)
        det = mr.cloud_inputs(spec, free)
        out = {"kind": kind, "lorentzian": lor, "logp": lp}
        for key, d in data.items():
            out[f"data_{key}_spectral"] = d.spectral
            out[f"data_{key}_brightness"] = d.brightness
            out[f"data_{key}_noise"] = d.noise
            out[f"mu_{key}"] = mu[key]
        for k, v in free.items():
            out[f"free_{k}"] = v
        for k, v in baseline.items():
            out[f"free_{k}"] = v
        for k, v in grads.items():
            out[f"grad_{k}"] = v
        for k, v in det.items():
            out[f"det_{k}"] = np.broadcast_to(v, (6, 3)).copy()
        name = f"{kind}_{'pv' if lor else 'gauss'}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, "finite draws:", int(np.isfinite(lp).sum()), "/ 6")


if __name__ == "__main__":
    main()
