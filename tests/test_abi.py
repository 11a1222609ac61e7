"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/lumcu.h declares; without a GPU it fails loudly instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest

import lum_b200
from lum_b200 import build as lbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lumcu_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    lbuild.build_all()
    return lum_b200.load_library()


def test_header_symbols_exported(lib):
    names = _declared("lumcu.h")
    assert len(names) >= 25
    raw = ctypes.CDLL(lum_b200.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/lumcu.h but not exported"
    # and the Python binding covers exactly the declared surface
    assert sorted(lum_b200.SYMBOLS) == names


def test_version_and_error_string(lib):
    assert lib.lumcu_version() >= 100
    assert isinstance(lib.lumcu_last_error(), bytes)


def test_argument_validation_without_device(lib):
    d = lum_b200.LumcuDesc((ctypes.c_int32 * 3)(48, 48, 16), 0, 64, 0.01, 0, 0)   # capacity 0
    ctx = ctypes.c_void_p()
    assert lib.lumcu_create(ctypes.byref(d), ctypes.byref(ctx)) == -1
    assert b"palette_capacity" in lib.lumcu_last_error()
    d = lum_b200.LumcuDesc((ctypes.c_int32 * 3)(48, 48, 16), 15, 48, 0.01, 0, 0)  # bad ray count
    assert lib.lumcu_create(ctypes.byref(d), ctypes.byref(ctx)) == -1
    assert lib.lumcu_update(None, 0, 1, 0) == -1
    assert lib.lumcu_destroy(None) == 0


def test_no_cpu_fallback(lib):
    """On a box without a CUDA device the stage must refuse to exist (no silent CPU path)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("CUDA device present")
    with pytest.raises(lum_b200.LumcuError, match="no CPU fallback"):
        lum_b200.RadianceStage((48, 48, 16), 15)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under lum_b200/ or include/ may reference it."""
    bad = []
    for base in ("lum_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".c")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"\b(import oracle|from oracle|lum_oracle|liblum_oracle)", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_api_mirror_symbols_exported():
    """liblum.so exports every lum_* / lum_b200_* function include/clum.h declares, and the ctypes binding
    (lum_b200/clum.py) covers exactly that surface."""
    from lum_b200 import clum
    lbuild.build_all()
    src = open(os.path.join(ROOT, "include", "clum.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(lum_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 37 and "lum_b200_glossy" in names
    raw = ctypes.CDLL(clum.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/clum.h but not exported"
    assert sorted(clum.SYMBOLS) == names
