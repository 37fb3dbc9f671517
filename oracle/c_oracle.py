"""ctypes binding of the plain-C oracle (oracle/chi_oracle.c): test infrastructure and the CPU
baseline of bench.py.  ``build()`` compiles it with gcc + OpenMP into oracle/_build/."""

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "chi_oracle.c")
LIB = os.path.join(HERE, "_build", "libchi_oracle.so")
_dp = C.POINTER(C.c_double)


class _Dataset(C.Structure):
    _fields_ = [("S", C.c_int), ("v", _dp), ("y", _dp), ("sig", _dp)]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.chi_oracle_logp_grad.restype = C.c_int
        lib.chi_oracle_logp_grad.argtypes = [C.c_int, C.c_int, C.POINTER(_Dataset), C.POINTER(_Dataset), C.c_double,
                                             _dp, _dp, _dp, C.c_int]
        _lib = lib
    return _lib


INPUTS = ("velocity", "fwhm2", "fwhm_L", "tau_total", "tspin", "filling_factor", "absorption_weight")


def logp_grad(data, det, bg_temp, n_threads=0):
    """``data``: {"emission"/"absorption": OracleData}; ``det``: dict of the seven likelihood
    inputs ``[B, N]``.  The constant baseline (``_brightness_offset``) is subtracted from the
    brightness.  Returns ``(logp [B], grad dict of [B, N], threads_used)``."""
    lib = _load()
    B, N = det["velocity"].shape
    block = np.ascontiguousarray(np.stack([np.broadcast_to(det[k], (B, N)) for k in INPUTS], axis=1), dtype=np.float64)
    keep = []

    def ds(key):
        d = _Dataset()
        if key in data:
            o = data[key]
            arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in
                    (o.spectral, o.brightness - o._brightness_offset, o.noise)]
            keep.extend(arrs)
            d.S = arrs[0].shape[0]
            d.v, d.y, d.sig = (a.ctypes.data_as(_dp) for a in arrs)
        return d

    de, da = ds("emission"), ds("absorption")
    logp = np.empty(B)
    grad = np.empty((B, 7, N))
    used = lib.chi_oracle_logp_grad(B, N, C.byref(de), C.byref(da), float(bg_temp), block.ctypes.data_as(_dp),
                                    logp.ctypes.data_as(_dp), grad.ctypes.data_as(_dp), int(n_threads))
    if used < 0:
        raise ValueError("too many clouds for the C oracle")
    return logp, {k: grad[:, i, :] for i, k in enumerate(INPUTS)}, used
