"""CPU: the C-ABI library builds, loads and exports every symbol include/b200dd.h declares;
host-only entry points work; compute entry points fail LOUDLY without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from blah2_b200 import capi
from oracle import blah2_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200dd.h")).read()
    return sorted(set(re.findall(r"B200DD_API[^;(]*?\b(b200dd_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 25
    assert sorted(n for n, _, _ in capi.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(capi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_host_only_entry_points():
    capi.load()
    for v in (0, 1, 104, 3322, 19043, 65534):
        assert capi.next_hamming(v) == O.next_hamming(v)
    assert capi.device_count() >= 0


def test_compute_fails_loudly_without_gpu():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    from blah2_b200.process import Ambiguity, WienerHopf, CfarDetector1D, SpectrumAnalyser
    with pytest.raises(capi.B200ddError) as e:
        Ambiguity(-10, 300, -300, 300, 2000000, 1000000)
    assert e.value.code == capi.ERR_CUDA
    with pytest.raises(capi.B200ddError) as e:
        SpectrumAnalyser(2000000, 2000.0)
    assert e.value.code == capi.ERR_CUDA
    with pytest.raises(capi.B200ddError):
        WienerHopf(-10, 400, 100000)
    with pytest.raises(capi.B200ddError):
        CfarDetector1D(1e-5, 2, 6, 5, 15.0)


def test_argument_validation_without_gpu():
    lib = capi.load()
    h = C.c_void_p()
    assert lib.b200dd_caf_create(None, C.byref(h)) == capi.ERR_ARG
    p = capi.CafParams(10, 5, -100, 100, 1000, 1000, 0, -1)   # delay_max < delay_min
    assert lib.b200dd_caf_create(C.byref(p), C.byref(h)) == capi.ERR_GEOMETRY
    assert lib.b200dd_wh_create(0, 0, 1000, -1, C.byref(h)) == capi.ERR_GEOMETRY     # zero taps
    assert lib.b200dd_wh_create(0, 5000, 100000, -1, C.byref(h)) == capi.ERR_GEOMETRY  # too many taps
    assert b"taps" in lib.b200dd_last_error()
    # SpectrumAnalyser: where the reference divides by zero (bandwidth > n, SpectrumAnalyser.cpp:17) or converts an
    # out-of-range double (bandwidth <= 0 / NaN, :16) the ABI answers ERR_GEOMETRY before touching a device
    assert lib.b200dd_spectrum_create(1000, 2000.0, -1, C.byref(h)) == capi.ERR_GEOMETRY
    assert lib.b200dd_spectrum_create(1000, -5.0, -1, C.byref(h)) == capi.ERR_GEOMETRY
    assert lib.b200dd_spectrum_create(1000, float("nan"), -1, C.byref(h)) == capi.ERR_GEOMETRY
    assert lib.b200dd_spectrum_create(1000, 100.0, -1, None) == capi.ERR_ARG


def test_product_never_imports_the_oracle():
    """The shipped package must not import / call anything under oracle/."""
    pkg = os.path.join(ROOT, "blah2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), os.path.join(dirpath, f)
