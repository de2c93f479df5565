"""GPU: SpectrumAnalyser (C ABI b200dd_spectrum_*, blah2_b200/csrc/spectrum.cu) against the golden fixtures
of the compiled reference (tests/golden/spectrum_*.npz), the numpy oracle, and -- at BASELINE sizes -- through
size-independent properties of the DFT (a single tone, linearity, the folded-sequence identity).

Tolerance: the kernels are FP64 end to end, inputs are read exactly -> 1e-11 relative to max |spectrum|
(the reference's own FP64 FFT differs from numpy's by ~5e-16 on the same inputs)."""
import os

import numpy as np
import pytest
import torch

from blah2_b200 import capi
from blah2_b200.process import Pipeline, SpectrumAnalyser
from blah2_b200.scene import make_scene, random_iq, Target
from oracle import blah2_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-11


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", ["spectrum_a", "spectrum_b", "spectrum_c", "spectrum_d"])
def test_cuda_spectrum_vs_reference_golden(name, relerr):
    d = gold(name)
    n, bw, seed = int(d["params"][0]), float(d["params"][1]), int(d["params"][2])
    x, _ = random_iq(n, seed)
    sa = SpectrumAnalyser(n, bw)
    assert (sa.decimation, sa.nSpectrum, sa.nfft) == O.spectrum_geometry(n, bw)
    spec, freq = sa.process(x)
    assert spec.shape == d["spectrum"].shape and freq.shape == d["frequency"].shape == (0,)
    e = relerr(spec, d["spectrum"])
    assert e[0] < TOL and e[1] < TOL, e


@pytest.mark.parametrize("n,bw", [(2000, 2000.0), (2001, 2000.0), (40000, 2000.0), (123457, 2000.0), (50000, 7.0),
                                  (65536, 9000.0), (30000, 12.5), (8, 8.0), (9, 2.0), (131072, 5000.0), (60000, 40000.0)])
def test_cuda_spectrum_vs_oracle_shapes(n, bw, relerr):
    """ragged sizes: nfft < n, decimation 1, one column tile / many, odd decimation, odd and even bin counts (the
    two fold kernels), more than 12288 bins (the DFT kernel then reads g from global memory), tiny inputs"""
    x, _ = random_iq(n, 31)
    sa = SpectrumAnalyser(n, bw)
    spec, _ = sa.process(x)
    ref, _ = O.spectrum_process(x, n, bw)
    e = relerr(spec, ref)
    assert e[0] < TOL and e[1] < TOL, (e, sa.decimation, sa.nSpectrum)


def test_device_float2_path_equals_host_path_on_int16_samples(relerr):
    """int16-valued IQ is exact in float32: the float2 device entry point must give the host path's result,
    twice in a row (no state carried between calls) and into a caller-provided buffer."""
    n, bw = 200000, 2000.0
    sc = make_scene(n, 2e6, seed=3, targets=[Target(20, 100.0, -30.0)])
    sa = SpectrumAnalyser(n, bw)
    host, _ = sa.process(sc.x)
    dx = torch.from_numpy(sc.x.astype(np.complex64)).cuda()
    sa.process_device(dx)
    dev = sa.fetch()
    # same samples, but the float2 kernel folds two columns per thread with its own row chunking: equal to rounding
    assert relerr(dev, host)[0] < 1e-13
    sa.process_device(dx)
    assert np.array_equal(sa.fetch(), dev)      # deterministic: every sum has a fixed order
    out = torch.empty(sa.nSpectrum, dtype=torch.complex128, device="cuda")
    st = torch.cuda.Stream()
    sa.process_device(dx, d_spectrum=out, stream=st.cuda_stream)
    st.synchronize()
    assert np.array_equal(out.cpu().numpy(), dev)
    # an input that starts on an odd float2 (8-byte aligned only) takes the one-column kernel: same answer
    buf = torch.empty(n + 1, dtype=torch.complex64, device="cuda")
    buf[1:].copy_(dx)
    sa.process_device(buf[1:])
    assert relerr(sa.fetch(), host)[0] < 1e-13
    ref, _ = O.spectrum_process(sc.x, n, bw)
    assert relerr(host, ref)[0] < TOL


def test_full_size_properties():
    """BASELINE sizes (2e6 and 2e7 samples, bandwidth 2000 as blah2.cpp:198), no oracle run needed:
    (1) a pure tone placed on a kept bin comes out as nfft in that bin and ~0 elsewhere;
    (2) linearity: S(a x1 + x2) = a S(x1) + S(x2);
    (3) the decimated spectrum equals the nSpectrum-point DFT of the phase-weighted, folded input, the fold done
        here in numpy FP64 (a reshape and one weighted sum) and the small DFT by numpy's FFT."""
    for n in (2000000, 20000000):
        bw = 2000.0
        dec, ns, nfft = O.spectrum_geometry(n, bw)
        sa = SpectrumAnalyser(n, bw)
        k0 = nfft // 2 + 1
        m0 = 777
        k = (m0 * dec + k0) % nfft
        ph = (np.arange(nfft, dtype=np.int64) * k) % nfft
        tone = np.exp(2j * np.pi * ph / nfft)
        s, _ = sa.process(tone)
        assert abs(s[m0] - nfft) < 1e-6 * nfft
        s[m0] = 0
        assert np.max(np.abs(s)) < 1e-6 * nfft
        rng = np.random.default_rng(5)
        x1 = rng.standard_normal(nfft) + 1j * rng.standard_normal(nfft)
        x2 = rng.standard_normal(nfft) + 1j * rng.standard_normal(nfft)
        a = 0.37 - 1.9j
        s1, _ = sa.process(x1)
        s2, _ = sa.process(x2)
        s12, _ = sa.process(a * x1 + x2)
        assert np.max(np.abs(s12 - (a * s1 + s2))) < 1e-10 * np.max(np.abs(s12))
        # (3) fold on the host in FP64 (a reshape + one weighted sum), then the small DFT by the oracle's FFT
        t = np.exp(-2j * np.pi * ((np.arange(nfft, dtype=np.int64) * k0) % nfft) / nfft)
        g = (x1 * t).reshape(dec, ns).sum(axis=0)
        ref = np.fft.fft(g)
        assert np.max(np.abs(s1 - ref)) < 1e-10 * np.max(np.abs(ref))


def test_geometry_limits_and_errors():
    lib = capi.load()
    import ctypes as C
    h = C.c_void_p()
    assert lib.b200dd_spectrum_create(1000, 2000.0, -1, C.byref(h)) == capi.ERR_GEOMETRY   # bandwidth > n: the reference divides by zero
    assert lib.b200dd_spectrum_create(1000, 0.0, -1, C.byref(h)) == capi.ERR_GEOMETRY
    assert lib.b200dd_spectrum_create(1000, float("nan"), -1, C.byref(h)) == capi.ERR_GEOMETRY
    assert lib.b200dd_spectrum_create(100000, 100000.0, -1, C.byref(h)) == capi.ERR_GEOMETRY   # 100000 bins > 65536
    sa = SpectrumAnalyser(4000, 100.0)
    with pytest.raises(capi.B200ddError):
        sa.process(np.zeros(100, dtype=np.complex128))   # fewer than nfft samples


def test_pipeline_spectrum_stage(relerr):
    """blah2.cpp:263-287 in one pipeline: spectrum of x first, then filter + CAF + detection; the other
    results must not change when the spectrum stage is enabled, on all three submit paths."""
    fs, n = 200000, 20000
    geom = (-5, 60, -500, 500, fs, n)
    det = dict(pfa=1e-4, nGuard=2, nTrain=6, minDelay=3, minDoppler=15.0, nCentroid=4)
    sc = make_scene(n, fs, seed=11, targets=[Target(17, 300.0, -25.0), Target(41, -200.0, -28.0)])
    base = Pipeline(*geom, roundHamming=True, clutter=(-5, 30), detection=det).process(sc.x, sc.y)
    pipe = Pipeline(*geom, roundHamming=True, clutter=(-5, 30), detection=det, spectrum_bandwidth=2000.0)
    ref, _ = O.spectrum_process(sc.x, n, 2000.0)
    assert pipe.n_spectrum == ref.shape[0]
    out = pipe.process(sc.x, sc.y)
    assert relerr(pipe.fetch_spectrum(), ref)[0] < TOL
    assert np.array_equal(out["map"], base["map"]) and out["noisePower"] == base["noisePower"]
    assert np.array_equal(out["detections"].delay, base["detections"].delay)
    # device path (float2) and int16 ingest path: the scene is int16-valued, so all paths agree
    dx = torch.from_numpy(sc.x.astype(np.complex64)).cuda()
    dy = torch.from_numpy(sc.y.astype(np.complex64)).cuda()
    pipe.submit_device(dx, dy)
    pipe.fetch()
    assert relerr(pipe.fetch_spectrum(), ref)[0] < TOL
    iq = np.empty((n, 4), dtype=np.int16)
    iq[:, 0], iq[:, 1], iq[:, 2], iq[:, 3] = sc.x.real, sc.x.imag, sc.y.real, sc.y.imag
    pipe.submit_host_rspduo(iq)
    pipe.fetch()
    assert relerr(pipe.fetch_spectrum(), ref)[0] < TOL
    # without the clutter filter the CAF consumes n_used < n samples but the spectrum still needs nfft
    p2 = Pipeline(*geom, roundHamming=True, spectrum_bandwidth=2000.0)
    p2.process(sc.x, sc.y)
    assert relerr(p2.fetch_spectrum(), ref)[0] < TOL
