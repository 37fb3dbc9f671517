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
