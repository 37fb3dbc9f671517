"""ctypes binding of ``libchi.so`` (C ABI declared in ``include/chi.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C
caribou_hi_b200/csrc``.  There is no CPU fallback: if the shared library is missing
or no sm_100 device is present, the product path raises.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CHI_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("CHI_LIB") or os.path.join(_HERE, "libchi.so")

# ---- constants mirrored from include/chi.h ------------------------------------
CHI_VERSION = 100
CHI_NSLOT = 10
CHI_NDET = 20
CHI_MAX_CLOUDS = 16
CHI_MAX_BASELINE = 8
CHI_COMM_ID_BYTES = 128

MODEL_PHYSICS = 0
MODEL_ABSORPTION = 1
MODEL_EMISSION = 2
MODEL_EMISSION_ABSORPTION = 3
MODEL_ABSORPTION_PHYSICAL = 4
MODEL_EMISSION_PHYSICAL = 5
MODEL_EMISSION_ABSORPTION_PHYSICAL = 6

OPT_SMALL_PATH, OPT_SMALL_NL, OPT_SMALL_CHUNKS, OPT_SMALL_STAMPS, OPT_NUTS_TICK, OPT_DEV_OVERLAP, OPT_EA_ONE_PASS, OPT_SMALL_CLUSTER = 0, 1, 2, 3, 4, 5, 6, 7

ERROR_NAMES = {0: "CHI_OK", -1: "CHI_EINVAL", -2: "CHI_ECUDA", -3: "CHI_ENODEV", -4: "CHI_ESTATE", -5: "CHI_ENOMEM"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_fp = C.POINTER(C.c_float)


class ChiConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32),
        ("n_clouds", C.c_int32),
        ("has_emission", C.c_int32),
        ("has_absorption", C.c_int32),
        ("lorentzian", C.c_int32),
        ("free_weight", C.c_int32),
        ("dtype", C.c_int32),
        ("reserved", C.c_int32),
        ("bg_temp", C.c_double),
        ("depth_nth_fwhm_power", C.c_double),
        ("prior_fwhm2", C.c_double),
        ("prior_velocity", C.c_double * 2),
        ("prior_log10_nHI", C.c_double * 2),
        ("prior_log10_n_alpha", C.c_double * 2),
        ("prior_fwhm_L", C.c_double),
        ("prior_amplitude", C.c_double),
    ]


class ChiDataset(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int32),
        ("n_baseline", C.c_int32),
        ("spectral", _dp),
        ("brightness", _dp),
        ("noise", _dp),
        ("basis", _dp),
        ("offset", _dp),
    ]


class ChiPriors(C.Structure):
    _fields_ = [
        ("beta", C.c_double * 2),
        ("nth_fwhm_1pc", C.c_double * 2),
        ("sigma_log10_NHI", C.c_double),
        ("baseline_sigma_e", C.c_double * 8),
        ("baseline_sigma_a", C.c_double * 8),
    ]


class ChiNutsConfig(C.Structure):
    _fields_ = [
        ("chains", C.c_int64),
        ("n_tune", C.c_int32),
        ("n_draws", C.c_int32),
        ("max_treedepth", C.c_int32),
        ("pool_mass", C.c_int32),
        ("seed", C.c_uint64),
        ("target_accept", C.c_double),
        ("init_step", C.c_double),
        ("lock_step", C.c_int32),
        ("reserved", C.c_int32),
    ]


class ChiError(RuntimeError):
    def __init__(self, code, text):
        self.code = code
        super().__init__(f"libchi {ERROR_NAMES.get(code, code)}: {text}")


# every symbol include/chi.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "chi_version": (C.c_int, []),
    "chi_device_count": (C.c_int, []),
    "chi_create": (C.c_int, [C.POINTER(_H), C.c_int, C.POINTER(ChiConfig)]),
    "chi_destroy": (None, [_H]),
    "chi_last_error": (C.c_char_p, [_H]),
    "chi_set_spectra": (C.c_int, [_H, C.c_int, C.POINTER(ChiDataset), C.POINTER(ChiDataset)]),
    "chi_stream": (C.c_void_p, [_H]),
    "chi_synchronize": (C.c_int, [_H]),
    "chi_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "chi_host_free": (C.c_int, [C.c_void_p]),
    "chi_spin_temp": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _dp]),
    "chi_spin_temp_vjp": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "chi_elementwise": (C.c_int, [_H, C.c_int, C.c_int64, _dp, _dp, _dp, _dp]),
    "chi_pseudo_voigt":(C.c_int, [_H, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]),
    "chi_radiative_transfer": (C.c_int, [_H, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, _dp]),
    "chi_pseudo_voigt_vjp": (C.c_int, [_H, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "chi_radiative_transfer_vjp": (C.c_int, [_H, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp, _dp]),
    "chi_logp_grad": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp]),
    "chi_logp_grad_dev": (C.c_int, [_H, C.c_int64] + [C.c_void_p] * 8),
    "chi_logp_grad_rows": (C.c_int, [_H, C.c_int64, C.c_int32, _ip, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp]),
    "chi_logp_grad_rows_dev": (C.c_int, [_H, C.c_int64, C.c_int32, _ip] + [C.c_void_p] * 8),
    "chi_logp_grad_rows_async": (C.c_int, [_H, C.c_int64, C.c_int32, _ip, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp,
                                           C.POINTER(C.c_int64)]),
    "chi_wait": (C.c_int, [_H, C.c_int64]),
    "chi_logp_sums_rows": (C.c_int, [_H, C.c_int64, C.c_int32, _ip, _dp, _dp, _dp, _ip, _dp, _dp]),
    "chi_logp_sums_rows_async": (C.c_int, [_H, C.c_int64, C.c_int32, _ip, _dp, _dp, _dp, _ip, _dp, _dp,
                                           C.POINTER(C.c_int64)]),
    "chi_stage_count": (C.c_int64, [_H]),
    "chi_spectra": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _dp, _dp]),
    "chi_spectra_dev": (C.c_int, [_H, C.c_int64] + [C.c_void_p] * 5),
    "chi_spectra_f32": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _fp, _fp]),
    "chi_spectra_f32_dev": (C.c_int, [_H, C.c_int64] + [C.c_void_p] * 5),
    "chi_predictive": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _ip, C.c_uint64, C.c_int64, _dp, _dp]),
    "chi_predictive_f32": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _ip, C.c_uint64, C.c_int64, _fp, _fp]),
    "chi_predictive_dev": (C.c_int, [_H, C.c_int64] + [C.c_void_p] * 4 + [C.c_uint64, C.c_int64] + [C.c_void_p] * 2),
    "chi_predictive_f32_dev": (C.c_int, [_H, C.c_int64] + [C.c_void_p] * 4 + [C.c_uint64, C.c_int64] + [C.c_void_p] * 2),
    "chi_spectra_vjp": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "chi_spectra_vjp_dev": (C.c_int, [_H, C.c_int64] + [C.c_void_p] * 8),
    "chi_deterministics": (C.c_int, [_H, C.c_int64, _dp, _dp]),
    "chi_set_priors": (C.c_int, [_H, C.POINTER(ChiPriors)]),
    "chi_model_logp_grad": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp]),
    "chi_model_logp_grad_dev": (C.c_int, [_H, C.c_int64] + [C.c_void_p] * 8),
    "chi_transform": (C.c_int, [_H, C.c_int64, C.c_int, _dp, _dp]),
    "chi_nuts_sample": (C.c_int, [_H, C.POINTER(ChiNutsConfig), _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _ip, _ip,
                                  _dp, _dp]),
    "chi_last_kernel_ms":(C.c_int, [_H, C.POINTER(C.c_float)]),
    "chi_stage_ms": (C.c_int, [_H, C.c_int, C.POINTER(C.c_float)]),
    "chi_launch_count": (C.c_int64, [_H]),
    "chi_set_option": (C.c_int, [_H, C.c_int, C.c_int64]),
    "chi_dev_join": (C.c_int, [_H]),
    "chi_comm_unique_id": (C.c_int, [C.c_void_p]),
    "chi_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    "chi_comm_destroy": (C.c_int, [_H]),
    "chi_comm_rank": (C.c_int, [_H]),
    "chi_comm_size": (C.c_int, [_H]),
    "chi_shard_sums_dev": (C.c_int, [_H, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "chi_allgather_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64]),
    "chi_allreduce_sum_dev": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "chi_comm_join": (C.c_int, [_H]),
    "chi_comm_synchronize": (C.c_int, [_H]),
    "chi_comm_stream": (C.c_void_p, [_H]),
    "chi_microbench_fp64": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "chi_microbench_fp32": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load ``libchi.so`` (once) and declare every prototype.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C caribou_hi_b200/csrc`. caribou_hi_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.chi_version() != CHI_VERSION:
        raise RuntimeError(f"libchi version {lib.chi_version()} != binding {CHI_VERSION}; rebuild")
    _lib = lib
    return lib


def _checked(a, dtype, ptr):
    # the library reads / writes through the raw pointer: a wrong dtype or a strided view would
    # corrupt host memory silently, so the marshalling layer refuses it
    if a is None:
        return None
    if a.dtype != dtype or not a.flags.c_contiguous:
        raise TypeError(f"libchi needs a C-contiguous {dtype} array, got {a.dtype} "
                        f"({'contiguous' if a.flags.c_contiguous else 'strided'}, shape {a.shape})")
    return a.ctypes.data_as(ptr)


def as_dp(a):
    """numpy float64 C-contiguous array -> double* (None -> NULL)."""
    return _checked(a, "float64", _dp)


def as_fp(a):
    return _checked(a, "float32", _fp)


def as_ip(a):
    return _checked(a, "int32", _ip)
