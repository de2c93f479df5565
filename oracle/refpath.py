"""oracle/refpath.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libblah2ref.so: the reference's UNMODIFIED hot-path
sources (compiled by oracle/Makefile against the FFTW3 / Armadillo shims in
oracle/shim/).  Used to (a) pin oracle/blah2_oracle.py, (b) generate tests/golden/,
(c) serve as the "reference" CPU baseline in bench.py.  /root/reference is only needed
to BUILD the library; the built .so travels to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libblah2ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
        _lib.refpath_next_hamming.restype = C.c_uint32
        _lib.refpath_next_hamming.argtypes = [C.c_uint32]
        _lib.refpath_chain_create.restype = C.c_void_p
        _lib.refpath_chain_create.argtypes = [C.c_int32] * 4 + [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int32,
                                                                C.c_int32, C.c_double, C.c_int, C.c_int, C.c_int,
                                                                C.c_double, C.c_uint32]
        _lib.refpath_chain_destroy.argtypes = [C.c_void_p]
        _lib.refpath_chain_run.restype = C.c_int
        _lib.refpath_chain_run.argtypes = [C.c_void_p] + [C.c_void_p] * 7 + [C.c_uint32, C.c_void_p]
        _lib.refpath_cfar.restype = C.c_uint32
        _lib.refpath_centroid.restype = C.c_uint32
        _lib.refpath_interpolate.restype = C.c_uint32
        _lib.refpath_last_error.restype = C.c_char_p
        _lib.refpath_spectrum_process.restype = C.c_int
        _lib.refpath_spectrum_process.argtypes = [C.c_uint32, C.c_double, C.c_void_p, C.c_uint32, C.c_void_p,
                                                  C.c_void_p, C.c_uint32, C.c_void_p]
    return _lib


def _check(rc, bad=(-1000,)):
    if rc in bad:
        raise RuntimeError("refpath: " + lib().refpath_last_error().decode("utf-8", "replace"))
    return rc


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c128(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.complex128))


def next_hamming(v: int) -> int:
    return int(lib().refpath_next_hamming(int(v)))


def ambiguity_geometry(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming=False) -> dict:
    nDel, nDop, nCorr, nfft = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    cpi, mid = C.c_double(), C.c_double()
    _check(lib().refpath_ambiguity_geometry(C.c_int32(delayMin), C.c_int32(delayMax), C.c_int32(dopplerMin),
                                     C.c_int32(dopplerMax), C.c_uint32(fs), C.c_uint32(n), C.c_int(int(roundHamming)),
                                     C.byref(nDel), C.byref(nDop), C.byref(nCorr), C.byref(nfft), C.byref(cpi),
                                     C.byref(mid)))
    return dict(nDelayBins=nDel.value, nDopplerBins=nDop.value, nCorr=nCorr.value, nfft=nfft.value, cpi=cpi.value,
                dopplerMiddle=mid.value)


def ambiguity_process(x, y, delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming=False) -> dict:
    g = ambiguity_geometry(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming)
    x, y = _c128(x), _c128(y)
    assert x.shape == y.shape
    nDop, nDel = g["nDopplerBins"], g["nDelayBins"]
    m = np.empty((nDop, nDel), dtype=np.complex128)
    delay = np.empty(nDel, dtype=np.int32)
    doppler = np.empty(nDop, dtype=np.float64)
    metrics = np.empty(2, dtype=np.float64)
    left = np.empty(2, dtype=np.uint32)
    _check(lib().refpath_ambiguity_process(C.c_int32(delayMin), C.c_int32(delayMax), C.c_int32(dopplerMin),
                                    C.c_int32(dopplerMax), C.c_uint32(fs), C.c_uint32(n), C.c_int(int(roundHamming)),
                                    _p(x), _p(y), C.c_uint32(x.shape[0]), _p(m), _p(delay), _p(doppler), _p(metrics),
                                    _p(left)))
    g.update(map=m, delay=delay, doppler=doppler, noisePower=float(metrics[0]), maxPower=float(metrics[1]),
             leftover=(int(left[0]), int(left[1])))
    return g


def wienerhopf_process(x, y, delayMin, delayMax):
    x, y = _c128(x), _c128(y).copy()
    ok = _check(lib().refpath_wienerhopf_process(C.c_int32(delayMin), C.c_int32(delayMax), C.c_uint32(x.shape[0]),
                                                 _p(x), _p(y)))
    return bool(ok), y


def spectrum_process(x, n, bandwidth, cap=1 << 17):
    """The reference's SpectrumAnalyser(n, bandwidth).process on an IqData holding all of x:
    (spectrum, frequency, samples left in x)."""
    x = _c128(x)
    spec = np.empty(cap, dtype=np.complex128)
    freq = np.empty(cap, dtype=np.float64)
    counts = np.zeros(3, dtype=np.uint32)
    _check(lib().refpath_spectrum_process(C.c_uint32(int(n)), C.c_double(float(bandwidth)), _p(x), x.shape[0],
                                          _p(spec), _p(freq), cap, _p(counts)))
    assert counts[0] <= cap and counts[1] <= cap
    return spec[:counts[0]].copy(), freq[:counts[1]].copy(), int(counts[2])


def set_metrics(m):
    m = _c128(m)
    out = np.empty(2, dtype=np.float64)
    lib().refpath_set_metrics(_p(m), C.c_uint32(m.shape[0]), C.c_uint32(m.shape[1]), _p(out))
    return float(out[0]), float(out[1])


def _det_bufs(cap):
    return [np.empty(cap, dtype=np.float64) for _ in range(3)]


def cfar_1d(m, delay, doppler, noisePower, pfa, nGuard, nTrain, minDelay, minDoppler):
    m = _c128(m)
    delay = np.ascontiguousarray(delay, dtype=np.int32)
    doppler = np.ascontiguousarray(doppler, dtype=np.float64)
    cap = m.size
    od, of, os_ = _det_bufs(cap)
    n = lib().refpath_cfar(C.c_double(pfa), C.c_int(nGuard), C.c_int(nTrain), C.c_int(minDelay),
                           C.c_double(minDoppler), _p(m), C.c_uint32(m.shape[0]), C.c_uint32(m.shape[1]), _p(delay),
                           _p(doppler), C.c_double(noisePower), _p(od), _p(of), _p(os_), C.c_uint32(cap))
    _check(n, (0xFFFFFFFF,))
    return od[:n].copy(), of[:n].copy(), os_[:n].copy()


def centroid(delay, doppler, snr, nDelay, nDoppler, resolutionDoppler):
    d = np.ascontiguousarray(delay, dtype=np.float64)
    f = np.ascontiguousarray(doppler, dtype=np.float64)
    s = np.ascontiguousarray(snr, dtype=np.float64)
    cap = max(1, d.shape[0])
    od, of, os_ = _det_bufs(cap)
    n = lib().refpath_centroid(C.c_uint32(nDelay), C.c_uint32(nDoppler), C.c_double(resolutionDoppler), _p(d), _p(f),
                               _p(s), C.c_uint32(d.shape[0]), _p(od), _p(of), _p(os_), C.c_uint32(cap))
    _check(n, (0xFFFFFFFF,))
    return od[:n].copy(), of[:n].copy(), os_[:n].copy()


def interpolate(delay, doppler, snr, m, mdelay, mdoppler, noisePower, doDelay=True, doDoppler=True):
    d = np.ascontiguousarray(delay, dtype=np.float64)
    f = np.ascontiguousarray(doppler, dtype=np.float64)
    s = np.ascontiguousarray(snr, dtype=np.float64)
    m = _c128(m)
    mdelay = np.ascontiguousarray(mdelay, dtype=np.int32)
    mdoppler = np.ascontiguousarray(mdoppler, dtype=np.float64)
    cap = max(1, d.shape[0])
    od, of, os_ = _det_bufs(cap)
    n = lib().refpath_interpolate(C.c_int(int(doDelay)), C.c_int(int(doDoppler)), _p(d), _p(f), _p(s),
                                  C.c_uint32(d.shape[0]), _p(m), C.c_uint32(m.shape[0]), C.c_uint32(m.shape[1]),
                                  _p(mdelay), _p(mdoppler), C.c_double(noisePower), _p(od), _p(of), _p(os_),
                                  C.c_uint32(cap))
    _check(n, (0xFFFFFFFF,))
    return od[:n].copy(), of[:n].copy(), os_[:n].copy()


class Chain:
    """Persistent reference objects (constructed once, as src/blah2.cpp:154-183 does)."""

    def __init__(self, delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming=True, clutter=None,
                 pfa=1e-5, nGuard=2, nTrain=6, minDelay=5, minDoppler=15.0, nCentroid=6):
        self.geom = ambiguity_geometry(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming)
        self.n = n
        cl = clutter if clutter is not None else (0, 0)
        self.h = lib().refpath_chain_create(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, int(roundHamming),
                                            int(clutter is not None), cl[0], cl[1], pfa, nGuard, nTrain, minDelay,
                                            minDoppler, nCentroid)
        if not self.h:
            _check(-1000)

    def run(self, x, y, want_map=True):
        x, y = _c128(x), _c128(y)
        assert x.shape[0] == self.n and y.shape[0] == self.n
        nDop, nDel = self.geom["nDopplerBins"], self.geom["nDelayBins"]
        m = np.empty((nDop, nDel), dtype=np.complex128) if want_map else None
        metrics = np.empty(2, dtype=np.float64)
        cap = nDop * nDel
        od, of, os_ = _det_bufs(cap)
        stage = np.zeros(3, dtype=np.float64)
        n = lib().refpath_chain_run(self.h, _p(x), _p(y), _p(m) if want_map else None, _p(metrics), _p(od), _p(of),
                                    _p(os_), C.c_uint32(cap), _p(stage))
        _check(n)
        if n < 0:
            return dict(skipped=True, stage_ms=stage)
        return dict(skipped=False, map=m, noisePower=float(metrics[0]), maxPower=float(metrics[1]),
                    detections=(od[:n].copy(), of[:n].copy(), os_[:n].copy()), stage_ms=stage)

    def close(self):
        if self.h:
            lib().refpath_chain_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
