"""CPU-only checks of the boundary: the library loads, exports every symbol
include/chi.h declares, and refuses to work without an sm_100 device."""

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "chi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chi_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from caribou_hi_b200 import _lib

    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in chi.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype"
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.chi_version() == _lib.CHI_VERSION


def test_header_constants_match_binding():
    from caribou_hi_b200 import _lib

    text = open(os.path.join(ROOT, "include", "chi.h")).read()
    for name in ("CHI_NSLOT", "CHI_NDET", "CHI_MAX_CLOUDS", "CHI_MAX_BASELINE", "CHI_VERSION"):
        m = re.search(rf"#define {name} (\d+)", text)
        assert m and int(m.group(1)) == getattr(_lib, name)
    enum = dict(re.findall(r"(CHI_MODEL_[A-Z_]+) = (\d+)", text))
    for k, v in enum.items():
        assert getattr(_lib, k.replace("CHI_", "")) == int(v)
    # every chi_set_option knob of the header has its constant in the binding and its name in Engine.set_option
    import inspect
    from caribou_hi_b200 import engine

    opts = dict(re.findall(r"#define (CHI_OPT_[A-Z_]+) (\d+)", text))
    count = int(opts.pop("CHI_OPT_COUNT"))
    assert len(opts) >= 8 and max(map(int, opts.values())) < count
    src = inspect.getsource(engine.Engine.set_option)
    for k, v in opts.items():
        assert getattr(_lib, k.replace("CHI_", "")) == int(v), k
        assert f"_lib.{k.replace('CHI_', '')}" in src, f"{k} is not reachable through Engine.set_option"


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from caribou_hi_b200 import Engine
    from caribou_hi_b200._lib import ChiError

    with pytest.raises(ChiError) as e:
        Engine("emission", 2)
    assert "no CPU fallback" in str(e.value) or "CHI_ENODEV" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "caribou_hi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_marshalling_refuses_wrong_dtype_and_strided_arrays():
    """The library writes through raw pointers: float32, strided or wrongly shaped output arrays must be
    refused on the Python side (silent host memory corruption otherwise)."""
    import numpy as np

    from caribou_hi_b200 import _lib
    from caribou_hi_b200.engine import _out_f64, _sightline

    assert _lib.as_dp(None) is None
    _lib.as_dp(np.zeros(4))
    with pytest.raises(TypeError):
        _lib.as_dp(np.zeros(4, dtype=np.float32))
    with pytest.raises(TypeError):
        _lib.as_dp(np.zeros((4, 4))[:, 1])
    with pytest.raises(TypeError):
        _lib.as_ip(np.zeros(4, dtype=np.int64))
    with pytest.raises(TypeError):
        _lib.as_fp(np.zeros(4))
    assert _out_f64(None, "logp", (3,)).shape == (3,)
    ok = np.empty((3, 2))
    assert _out_f64({"grad": ok}, "grad", (3, 2)) is ok
    for bad in (np.empty((3, 2), dtype=np.float32), np.empty((4, 2)), np.empty((3, 4))[:, ::2]):
        with pytest.raises(ValueError):
            _out_f64({"grad": bad}, "grad", (3, 2))
    with pytest.raises(ValueError):
        _sightline(np.zeros(2), 3)
    assert _sightline([0, 1, 2], 3).dtype == np.int32
