"""CPU: the PRODUCT's host-side geometry / plan code (b200dd_caf_plan, b200dd_spectrum_plan in libb200dd.so -- the
host half of the create calls, no device needed) against the reference's known answers, the golden fixtures written
by the compiled reference (tests/golden/geometry.npz, caf_*.npz axes, spectrum_*.npz) and the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from blah2_b200 import capi
from oracle import blah2_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_constructor_known_answers_like_testambiguity_cpp():
    g, delay, doppler = capi.caf_plan(-10, 300, -300, 300, 2000000, 1000000, False)   # TestAmbiguity.cpp:73-93
    assert (g.n_corr, g.n_delay_bins, g.n_doppler_bins, g.nfft, g.doppler_middle) == (3322, 311, 301, 6643, 0)
    assert abs(g.cpi - 0.5) < 0.02
    assert capi.caf_plan(-10, 300, -300, 300, 2000000, 1000000, True)[0].nfft == 6750   # :110-115
    assert delay[0] == -10 and delay[-1] == 300 and doppler[150] == 0.0


def test_geometry_matches_the_compiled_reference_golden():
    for row in gold("geometry")["rows"]:
        args = [int(v) for v in row[:6]] + [bool(row[6])]
        g, _, _ = capi.caf_plan(*args)
        assert (g.n_delay_bins, g.n_doppler_bins, g.n_corr, g.nfft) == tuple(int(v) for v in row[7:11]), args
        assert g.cpi == row[11] and g.doppler_middle == row[12]
        assert g.n_used == g.n_doppler_bins * g.n_corr
        # the plan: segments cover a batch, the hop leaves room for every wanted lag, groups / parts fit the segments
        assert g.range_segments * g.range_hop >= g.n_corr
        assert g.range_hop + g.n_delay_bins - 1 <= g.range_fft_len
        assert 1 <= g.range_groups * g.range_parts <= max(1, g.range_segments)
        assert g.doppler_fft_len >= 2 * g.n_doppler_bins - 1


@pytest.mark.parametrize("name", ["caf_a", "caf_b", "caf_c", "chain_a"])
def test_axes_are_bit_identical_to_the_reference_maps(name):
    """Interpolate matches Doppler values with == (Map.cpp:103-113): the axes must be bit-exact."""
    d = gold(name)
    geom = [int(v) for v in d["geom"][:6]] + [bool(d["geom"][6])]
    _, delay, doppler = capi.caf_plan(*geom)
    assert np.array_equal(delay, d["delay"]) and np.array_equal(doppler, d["doppler"])


def test_survey_size_table_and_default_plan():
    # SURVEY.md s8 size table; the plan the round's measurements were made with (148 SMs assumed without a device)
    g = capi.caf_plan(0, 299, -128, 128, 2000000, 2000000, True)[0]
    assert (g.n_delay_bins, g.n_doppler_bins, g.n_corr, g.nfft) == (300, 257, 7782, 15625)
    # round 2: the DIT range kernel has no warp groups; a batch is split into parts until ~5 CTAs per SM exist
    assert (g.range_fft_len, g.range_segments, g.range_groups, g.range_parts, g.doppler_fft_len) == (2048, 5, 1, 2, 1024)
    g = capi.caf_plan(0, 511, -256, 256, 10000000, 20000000, True)[0]
    assert (g.n_delay_bins, g.n_doppler_bins, g.n_corr, g.nfft) == (512, 1025, 19512, 39366)
    assert (g.range_fft_len, g.range_groups, g.range_parts, g.doppler_fft_len) == (2048, 1, 1, 4096)
    g = capi.caf_plan(0, 511, -512, 512, 20000000, 80000000, True)[0]
    assert (g.n_doppler_bins, g.n_corr) == (4097, 19526)


def test_plan_rejects_what_create_rejects():
    lib = capi.load()
    g = capi.CafGeometry()
    p = capi.CafParams(10, 5, -100, 100, 1000, 1000, 0, -1)
    assert lib.b200dd_caf_plan(C.byref(p), C.byref(g), None, 0, None, 0) == capi.ERR_GEOMETRY
    p = capi.CafParams(0, 10, -100, 100, 0, 1000, 0, -1)
    assert lib.b200dd_caf_plan(C.byref(p), C.byref(g), None, 0, None, 0) == capi.ERR_ARG
    p = capi.CafParams(0, 20000, -100, 100, 100000, 100000, 0, -1)     # 20001 delay bins: no FFT plan
    assert lib.b200dd_caf_plan(C.byref(p), C.byref(g), None, 0, None, 0) == capi.ERR_GEOMETRY
    p = capi.CafParams(0, 10, -100, 100, 1000, 1000, 0, -1)
    small = np.empty(2, dtype=np.int32)
    assert lib.b200dd_caf_plan(C.byref(p), C.byref(g), capi.ptr(small), 2, None, 0) == capi.ERR_ARG   # capacity
    assert lib.b200dd_caf_plan(None, C.byref(g), None, 0, None, 0) == capi.ERR_ARG


@pytest.mark.parametrize("n,bw", [(2000000, 2000.0), (20000000, 2000.0), (5003, 97.0), (3999, 2000.0), (20000, 333.3),
                                  (6000, 2000.0), (8, 8.0)])
def test_spectrum_plan_matches_oracle_and_golden(n, bw):
    g, f = capi.spectrum_plan(n, bw)
    assert (g.decimation, g.n_spectrum, g.nfft) == O.spectrum_geometry(n, bw)
    assert g.n_frequency == 0 and f.shape == (0,) and O.spectrum_frequency(n, bw).shape == (0,)
    assert g.fold_chunks * g.fold_rows_per_chunk >= g.decimation
    for name in ("spectrum_a", "spectrum_b", "spectrum_c", "spectrum_d"):
        d = gold(name)
        if int(d["params"][0]) == n and float(d["params"][1]) == bw:
            assert d["spectrum"].shape == (g.n_spectrum,) and d["frequency"].shape == (g.n_frequency,)


def test_spectrum_plan_fences_the_reference_undefined_cases():
    lib = capi.load()
    g = capi.SpectrumGeometry()
    for n, bw in ((1000, 2000.0), (1000, 0.0), (1000, -1.0), (1000, float("nan")), (0, 10.0), (100000, 100000.0)):
        assert lib.b200dd_spectrum_plan(n, bw, C.byref(g), None, 0) == capi.ERR_GEOMETRY, (n, bw)
    assert lib.b200dd_spectrum_plan(1000, 10.0, None, None, 0) == capi.ERR_ARG
