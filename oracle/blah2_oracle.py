"""oracle/blah2_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU (numpy, float64) restatement of the reference's delay-Doppler hot path.  It is the
CHECKER for the CUDA product path; nothing under ``blah2_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The restatement is pinned (tests/test_oracle_*.py) against

  * the reference's own known-answer tests (TestAmbiguity.cpp:87-92,110-115;
    TestHammingNumber.cpp:15-17), and
  * outputs of the reference's UNMODIFIED sources compiled here into
    oracle/_ref/libblah2ref.so (oracle/Makefile), committed as fixtures under
    tests/golden/ by oracle/gen_golden.py.

Third-party arithmetic that is not under /root/reference:
  * FFTW3 (apt libfftw3-dev, unpinned, Dockerfile:12): an exact DFT -> numpy's
    pocketfft in float64 is an equivalent.
  * Armadillo 12.0.1 (lib/vcpkg.json:9,18) -> LAPACK zpotrf / ztrtrs: restated with
    scipy.linalg.cholesky / solve_triangular (the same LAPACK routines).  No reference
    test touches WienerHopf: PARITY UNPINNED by reference fixtures at that boundary;
    it is pinned to the compiled reference source + oracle/shim/armadillo only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

import os

try:  # scipy is present in the image; keep the import local to the solve
    import scipy.linalg as _sla
except Exception:  # pragma: no cover
    _sla = None
try:
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None


class _FFT:
    """The DFTs of this file.  Default: numpy's pocketfft on one thread.  BLAH2_ORACLE_WORKERS=k (k = -1: every
    core) switches to scipy.fft -- the same pocketfft algorithm -- with k worker threads, for the full-size
    BASELINE configurations (2e7 .. 8e7 samples) the GPU tests check live; results agree to ~1e-16."""

    @staticmethod
    def _workers():
        try:
            return int(os.environ.get("BLAH2_ORACLE_WORKERS", "0"))
        except ValueError:
            return 0

    def fft(self, a, n=None, axis=-1):
        w = self._workers()
        if w and _sfft is not None:
            return _sfft.fft(a, n=n, axis=axis, workers=w)
        return np.fft.fft(a, n=n, axis=axis)

    def ifft(self, a, n=None, axis=-1):
        w = self._workers()
        if w and _sfft is not None:
            return _sfft.ifft(a, n=n, axis=axis, workers=w)
        return np.fft.ifft(a, n=n, axis=axis)


_fft = _FFT()


# --------------------------------------------------------------------------------------
# HammingNumber  (src/process/meta/HammingNumber.cpp:38-48)
# --------------------------------------------------------------------------------------
def next_hamming(value: int) -> int:
    """First 5-smooth number STRICTLY greater than ``value`` (HammingNumber.cpp:38-48:
    the generator yields 1,2,3,4,5,6,8,... and returns the first ``i > value``)."""
    best = None
    p2 = 1
    while p2 <= 2 * (value + 1):
        p3 = p2
        while p3 <= 2 * (value + 1):
            p5 = p3
            while p5 <= 2 * (value + 1):
                if p5 > value and (best is None or p5 < best):
                    best = p5
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return int(best)


# --------------------------------------------------------------------------------------
# SpectrumAnalyser  (src/process/spectrum/SpectrumAnalyser.cpp:9-74)
# --------------------------------------------------------------------------------------
def spectrum_geometry(n: int, bandwidth: float):
    """(decimation, nSpectrum, nfft) as the constructor computes them (SpectrumAnalyser.cpp:16-18):
    ``decimation = n/bandwidth`` is a double division truncated into a uint32_t member, the other
    two are uint32_t integer arithmetic."""
    decimation = int(float(n) / float(bandwidth)) & 0xFFFFFFFF
    nSpectrum = (n // decimation) & 0xFFFFFFFF
    nfft = (nSpectrum * decimation) & 0xFFFFFFFF
    return decimation, nSpectrum, nfft


def spectrum_frequency(n: int, bandwidth: float) -> np.ndarray:
    """The vector handed to IqData::update_frequency (SpectrumAnalyser.cpp:57-67).  The loop counter
    ``i`` is a uint32_t (:34), so ``i = -nSpectrum/2`` is (2^32 - nSpectrum) / 2 and ``i < nSpectrum/2``
    is false at once for every nSpectrum < 2^31: the reference publishes an EMPTY frequency vector
    (confirmed on the compiled reference, tests/golden/spectrum_*.npz).  Restated literally."""
    decimation, nSpectrum, _ = spectrum_geometry(n, bandwidth)
    offset = bandwidth / 2 if decimation % 2 == 0 else 0.0
    start = ((-nSpectrum) & 0xFFFFFFFF) // 2
    stop = nSpectrum // 2
    i = np.arange(start, stop, dtype=np.float64) if start < stop else np.zeros(0)
    return ((i * bandwidth) + offset + 204640000) / 1000


def spectrum_process(x: np.ndarray, n: int, bandwidth: float):
    """SpectrumAnalyser::process (SpectrumAnalyser.cpp:31-74): nfft-point forward DFT of the first nfft
    samples (:36-40), ``fftshift[i] = X[(i + int(nfft/2) + 1) % nfft]`` (:43-47), every decimation-th
    entry kept (:50-54).  Returns (spectrum, frequency); x is read, not consumed (get_data copies, :35)."""
    decimation, nSpectrum, nfft = spectrum_geometry(n, bandwidth)
    X = _fft.fft(np.asarray(x, dtype=np.complex128)[:nfft])
    i = np.arange(0, nfft, decimation, dtype=np.int64)
    return X[(i + nfft // 2 + 1) % nfft], spectrum_frequency(n, bandwidth)


# --------------------------------------------------------------------------------------
# Ambiguity constructor  (src/process/ambiguity/Ambiguity.cpp:11-82)
# --------------------------------------------------------------------------------------
@dataclass
class Geometry:
    delayMin: int
    delayMax: int
    dopplerMin: int
    dopplerMax: int
    fs: int
    n: int
    roundHamming: bool
    nDelayBins: int = 0
    dopplerMiddle: float = 0.0
    nDopplerBins: int = 0
    nCorr: int = 0
    cpi: float = 0.0
    nfft: int = 0
    delay: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    doppler: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))

    @property
    def n_used(self) -> int:
        return self.nDopplerBins * self.nCorr


def ambiguity_geometry(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming=False) -> Geometry:
    g = Geometry(int(delayMin), int(delayMax), int(dopplerMin), int(dopplerMax), int(fs), int(n), bool(roundHamming))
    # Ambiguity.cpp:22  nDelayBins = static_cast<uint16_t>(delayMax - delayMin + 1)
    g.nDelayBins = (g.delayMax - g.delayMin + 1) & 0xFFFF
    # :23
    g.dopplerMiddle = (g.dopplerMin + g.dopplerMax) / 2.0
    # :26-36  count bins with the NOMINAL resolution fs/n
    res = 1.0 / (float(g.n) / float(g.fs))
    count = 1
    i = 1
    while g.dopplerMiddle + (i * res) <= g.dopplerMax:
        count += 2
        i += 1
    g.nDopplerBins = count & 0xFFFF  # uint16_t member (Ambiguity.h:86)
    # :39-40  nCorr = n / nDopplerBins into a uint16_t (Ambiguity.h:89)
    g.nCorr = (g.n // g.nDopplerBins) & 0xFFFF
    g.cpi = (float(g.nCorr) * g.nDopplerBins) / g.fs
    # :43-59  axes use the TRUE cpi
    res = 1.0 / g.cpi
    g.delay = np.arange(g.delayMin, g.delayMin + g.nDelayBins, dtype=np.int32)
    half = (g.nDopplerBins - 1) // 2
    ax = [g.dopplerMiddle]
    k = 1
    while len(ax) < g.nDopplerBins:
        ax.append(g.dopplerMiddle + (k * res))
        ax.insert(0, g.dopplerMiddle - (k * res))
        k += 1
    g.doppler = np.asarray(ax, dtype=np.float64)
    assert half * 2 + 1 == g.nDopplerBins
    # :62-65
    g.nfft = 2 * g.nCorr - 1
    if g.roundHamming:
        g.nfft = next_hamming(g.nfft)
    return g


# --------------------------------------------------------------------------------------
# Ambiguity::process  (src/process/ambiguity/Ambiguity.cpp:92-172)
# --------------------------------------------------------------------------------------
def ambiguity_prerotate(x: np.ndarray, g: Geometry) -> np.ndarray:
    """A2 -- Ambiguity.cpp:95-102: if dopplerMiddle != 0 every queued reference sample i
    is multiplied by exp(+j 2 pi dopplerMiddle i / fs)."""
    if g.dopplerMiddle == 0:
        return x
    i = np.arange(x.shape[0], dtype=np.float64)
    return x * np.exp(1j * 2.0 * np.pi * g.dopplerMiddle * (i / g.fs))


def range_matrix(x: np.ndarray, y: np.ndarray, g: Geometry, chunk: int = 64) -> np.ndarray:
    """A3+A4 -- Ambiguity.cpp:106-149.  Row i = lags delayMin..delayMax of the linear
    cross-correlation of batch i (samples outside the batch are zero because of the
    per-batch zero padding at :114-118):  R[i][j] = sum_n y_i[n+l] conj(x_i[n]),
    l = delayMin + j.  Computed exactly as the reference does: nfft-point FFTs of the
    zero-padded batches, Y conj(X) / nfft, unnormalised inverse, lag pick :132-146."""
    nD, nC, nL, nfft = g.nDopplerBins, g.nCorr, g.nDelayBins, g.nfft
    xb = x[: nD * nC].reshape(nD, nC)
    yb = y[: nD * nC].reshape(nD, nC)
    R = np.empty((nD, nL), dtype=np.complex128)
    lags = g.delayMin + np.arange(nL)
    # dataCorr[nDel + delayMin + j]  with dataCorr = [z[nfft-nDel .. nfft-1], z[0 .. nDel]]
    idx = np.where(lags >= 0, lags, nfft + lags)
    for s in range(0, nD, chunk):
        X = _fft.fft(xb[s : s + chunk], n=nfft, axis=1)
        Y = _fft.fft(yb[s : s + chunk], n=nfft, axis=1)
        Z = (Y * np.conj(X)) / float(nfft)
        z = _fft.ifft(Z, axis=1) * nfft  # FFTW backward is unnormalised
        R[s : s + chunk] = z[:, idx]
    return R


def doppler_transform(R: np.ndarray, g: Geometry) -> np.ndarray:
    """A5 -- Ambiguity.cpp:152-169: forward DFT of length nDop down every delay column,
    then out[k] = D[(k + nDop/2 + 1) % nDop] (an fftshift for odd nDop)."""
    nD = g.nDopplerBins
    D = _fft.fft(R, axis=0)
    k = (np.arange(nD) + nD // 2 + 1) % nD
    return D[k, :]


def ambiguity_process(x: np.ndarray, y: np.ndarray, g: Geometry):
    """Full Ambiguity::process.  Returns (map [nDop][nDel] complex128, leftover_x, leftover_y)
    -- the reference consumes nDop*nCorr samples from each FIFO (:108-112)."""
    x = np.asarray(x, dtype=np.complex128)
    y = np.asarray(y, dtype=np.complex128)
    x = ambiguity_prerotate(x, g)
    R = range_matrix(x, y, g)
    return doppler_transform(R, g), x.shape[0] - g.n_used, y.shape[0] - g.n_used


def range_matrix_direct(x: np.ndarray, y: np.ndarray, g: Geometry) -> np.ndarray:
    """Time-domain statement of A4 (independent of any FFT) for SMALL cases only."""
    nD, nC, nL = g.nDopplerBins, g.nCorr, g.nDelayBins
    R = np.zeros((nD, nL), dtype=np.complex128)
    for i in range(nD):
        xi = x[i * nC : (i + 1) * nC]
        yi = y[i * nC : (i + 1) * nC]
        for j in range(nL):
            l = g.delayMin + j
            if l >= 0:
                R[i, j] = np.sum(yi[l:] * np.conj(xi[: nC - l])) if l < nC else 0
            else:
                R[i, j] = np.sum(yi[: nC + l] * np.conj(xi[-l:])) if -l < nC else 0
    return R


# --------------------------------------------------------------------------------------
# Map::set_metrics  (src/data/Map.cpp:188-206)
# --------------------------------------------------------------------------------------
def set_metrics(m: np.ndarray):
    """noisePower = mean(10 log10 |z|) over all cells; maxPower = max(0, max value) - noise
    (the running max is initialised at 0, Map.cpp:193)."""
    with np.errstate(divide="ignore"):
        v = 10.0 * np.log10(np.abs(m))
    noise = float(np.sum(v) / (m.shape[0] * m.shape[1]))
    mx = max(0.0, float(np.max(v)))
    return noise, mx - noise


# --------------------------------------------------------------------------------------
# WienerHopf::process  (src/process/clutter/WienerHopf.cpp:58-163)
# --------------------------------------------------------------------------------------
def wienerhopf_weights(x: np.ndarray, y: np.ndarray, delayMin: int, delayMax: int):
    """W2-W4.  Returns (ok, w, a, b, xs).  nBins = delayMax - delayMin (NO +1, :12)."""
    x = np.asarray(x, dtype=np.complex128)
    y = np.asarray(y, dtype=np.complex128)
    N = x.shape[0]
    nBins = int(delayMax) - int(delayMin)
    # :65-69  dataX[i] = xData[(((i - delayMin) % nSamples) + nSamples) % nSamples] with i uint32_t and
    # delayMin int32_t: the subtraction is done in uint32 (wraps modulo 2^32).  For delayMin <= 0 that
    # is the circular shift x[(i + |delayMin|) mod N]; for delayMin > 0 the first delayMin samples
    # come from index (2^32 - (delayMin - i)) mod N -- reproduced literally.
    t = (np.arange(N, dtype=np.int64) - int(delayMin)) % (1 << 32)
    xs = x[((t % N) + N) % N]
    ys = y
    X = _fft.fft(xs)  # :72
    Y = _fft.fft(ys)  # :73
    # :76-84  a[k] = conj(IFFT_unnorm(|X|^2)[k]) / N
    dataA = _fft.ifft(X * np.conj(X)) * N
    a = np.conj(dataA[:nBins]) / float(N)
    # :100-108  b[k] = IFFT_unnorm(Y conj X)[k] / N
    dataB = _fft.ifft(Y * np.conj(X)) * N
    b = dataB[:nBins] / float(N)
    # :85-97  A = toeplitz(a) (no conjugation), then conj where i > j
    ii, jj = np.meshgrid(np.arange(nBins), np.arange(nBins), indexing="ij")
    A = a[np.abs(ii - jj)]
    A = np.where(ii > jj, np.conj(A), A)
    # :111-122  upper Cholesky A = R^H R, w = R^-1 (R^H)^-1 b ; failure -> return false
    try:
        Rm = _sla.cholesky(A, lower=False, check_finite=True)
    except Exception:
        return False, None, a, b, xs
    z = _sla.solve_triangular(Rm, b, trans="C", lower=False)
    w = _sla.solve_triangular(Rm, z, trans="N", lower=False)
    if not np.all(np.isfinite(w)):
        return False, None, a, b, xs
    return True, w, a, b, xs


def wienerhopf_apply(xs: np.ndarray, y: np.ndarray, w: np.ndarray) -> np.ndarray:
    """W5 -- :125-160: LINEAR convolution of w (nBins taps) with the shifted reference
    (zero history), via FFTs of the awkward length M = N + nBins + 1 in the reference;
    restated in the time domain through an FFT convolution of a 5-smooth length (the
    result is the same linear convolution)."""
    N = xs.shape[0]
    nBins = w.shape[0]
    L = N + nBins + 1
    Lf = next_hamming(L)
    F = _fft.ifft(_fft.fft(xs, n=Lf) * _fft.fft(w, n=Lf))
    return np.asarray(y, dtype=np.complex128) - F[:N]


def wienerhopf_process(x: np.ndarray, y: np.ndarray, delayMin: int, delayMax: int):
    """Returns (ok, y_filtered).  ok False <=> reference returns false and leaves y untouched."""
    ok, w, a, b, xs = wienerhopf_weights(x, y, delayMin, delayMax)
    if not ok:
        return False, np.asarray(y, dtype=np.complex128)
    return True, wienerhopf_apply(xs, y, w)


# --------------------------------------------------------------------------------------
# CfarDetector1D::process  (src/process/detection/CfarDetector1D.cpp:23-100)
# --------------------------------------------------------------------------------------
def _int8(v: int) -> int:
    v = int(v) & 0xFF
    return v - 256 if v >= 128 else v


def cfar_1d(m: np.ndarray, delay: np.ndarray, doppler: np.ndarray, noisePower: float, pfa: float, nGuard: int,
            nTrain: int, minDelay: int, minDoppler: float):
    """Returns (delay, doppler, snr) float64 arrays in the reference's emission order
    (row-major over Doppler rows then delay bins)."""
    nGuard, nTrain, minDelay = _int8(nGuard), _int8(nTrain), _int8(minDelay)
    nDop, nDel = m.shape
    o_delay, o_doppler, o_snr = [], [], []
    for i in range(nDop):
        if abs(doppler[i]) < minDoppler:  # :40
            continue
        row = m[i]
        sq = np.abs(row * row)  # :47  abs(z*z)
        with np.errstate(divide="ignore"):
            snr = 10.0 * np.log10(np.abs(row)) - noisePower  # :48
        for j in range(nDel):
            if delay[j] < minDelay:  # :53
                continue
            idx = [k for k in range(j - nGuard - nTrain, j - nGuard) if 0 < k < nDel]  # :59-64 (k > 0 !)
            idx += [k for k in range(j + nGuard + 1, j + nGuard + nTrain + 1) if 0 <= k < nDel]  # :66-71
            nCells = len(idx)
            if nCells == 0:
                continue  # alpha = 0 * (inf - 1) = NaN -> comparison false (:76-86)
            alpha = nCells * (math.pow(pfa, -1.0 / nCells) - 1)  # :76
            trainNoise = 0.0
            for k in idx:  # sequential sum, :78-81
                trainNoise += sq[k]
            trainNoise /= nCells
            if sq[j] > alpha * trainNoise:  # :86
                o_delay.append(float(j + delay[0]))  # :88
                o_doppler.append(float(doppler[i]))
                o_snr.append(float(snr[j]))
    return np.asarray(o_delay), np.asarray(o_doppler), np.asarray(o_snr)


# --------------------------------------------------------------------------------------
# Centroid::process  (src/process/detection/Centroid.cpp:19-73)
# --------------------------------------------------------------------------------------
def centroid(delay, doppler, snr, nDelay: int, nDoppler: int, resolutionDoppler: float):
    nDelay &= 0xFFFF
    nDoppler &= 0xFFFF
    n = len(snr)
    keep = []
    for i in range(n):
        # :28,34-35  uint16_t delayMin/delayMax = (int)delay[i] -/+ nDelay  (wraps mod 2^16)
        dmin = (int(delay[i]) - nDelay) & 0xFFFF
        dmax = (int(delay[i]) + nDelay) & 0xFFFF
        fmin = doppler[i] - (nDoppler * resolutionDoppler)
        fmax = doppler[i] + (nDoppler * resolutionDoppler)
        is_c = True
        for j in range(n):
            if j == i:
                continue
            if delay[j] > dmin and delay[j] < dmax and doppler[j] > fmin and doppler[j] < fmax:
                if snr[i] < snr[j]:
                    is_c = False
                    break
        if is_c:
            keep.append(i)
    keep = np.asarray(keep, dtype=np.int64)
    return np.asarray(delay)[keep], np.asarray(doppler)[keep], np.asarray(snr)[keep]


# --------------------------------------------------------------------------------------
# Interpolate::process  (src/process/detection/Interpolate.cpp:20-91)
# --------------------------------------------------------------------------------------
def _hz_to_bin(doppler_axis, hz) -> int:
    """Map::doppler_hz_to_bin, Map.cpp:103-113: exact == match, 0 when absent."""
    hit = np.nonzero(doppler_axis == hz)[0]
    return int(hit[0]) if hit.size else 0


def interpolate(delay, doppler, snr, m: np.ndarray, mdelay, mdoppler, noisePower: float, doDelay=True,
                doDoppler=True):
    def db(r, c):
        with np.errstate(divide="ignore"):
            return 10.0 * math.log10(abs(m[r][c])) - noisePower if abs(m[r][c]) > 0 else -math.inf

    od, of, os_ = [], [], []
    for i in range(len(snr)):
        intDelay, intDoppler = delay[i], doppler[i]
        intSnrDelay = snr[i]
        intSnrDoppler = snr[i]  # never updated: Interpolate.cpp:80 assigns intSnrDelay again
        if doDelay:
            if delay[i] == mdelay[0] or delay[i] == mdelay[-1]:  # :46-49
                continue
            r = _hz_to_bin(mdoppler, doppler[i])
            c = int(delay[i] - mdelay[0])
            s0, s1, s2 = db(r, c - 1), db(r, c), db(r, c + 1)  # :50-52
            if s1 < s0 or s1 < s2:  # :54-58
                continue
            intDelay = (s0 - s2) / (2 * (s0 - (2 * s1) + s2))  # :59
            intSnrDelay = s1 - (((s0 - s2) * intDelay) / 4)  # :60
            intDelay = delay[i] + intDelay  # :61
        if doDoppler:
            if doppler[i] == mdoppler[0] or doppler[i] == mdoppler[-1]:  # :67-70
                continue
            r = _hz_to_bin(mdoppler, doppler[i])
            c = int(delay[i] - mdelay[0])
            s0, s1, s2 = db(r - 1, c), db(r, c), db(r + 1, c)  # :71-73
            if s1 < s0 or s1 < s2:  # :75-78
                continue
            intDoppler = (s0 - s2) / (2 * (s0 - (2 * s1) + s2))  # :79
            intSnrDelay = s1 - (((s0 - s2) * intDoppler) / 4)  # :80 (sic: intSnrDelay)
            intDoppler = doppler[i] + ((mdoppler[1] - mdoppler[0]) * intDoppler)  # :81
        od.append(intDelay)
        of.append(intDoppler)
        os_.append(max(max(intSnrDelay, intSnrDoppler), snr[i]))  # :86
    return np.asarray(od, dtype=np.float64), np.asarray(of, dtype=np.float64), np.asarray(os_, dtype=np.float64)


# --------------------------------------------------------------------------------------
# whole chain in the order of src/blah2.cpp:268-287
# --------------------------------------------------------------------------------------
def chain(x, y, g: Geometry, clutter=None, det=None):
    """clutter = (delayMinClutter, delayMaxClutter) or None; det = dict(pfa,nGuard,nTrain,
    minDelay,minDoppler,nCentroid) or None.  Returns dict."""
    out = {"skipped": False}
    if clutter is not None:
        ok, y = wienerhopf_process(x, y, clutter[0], clutter[1])
        if not ok:
            out["skipped"] = True  # blah2.cpp:270-273 `continue`
            return out
    m, _, _ = ambiguity_process(x, y, g)
    noise, mx = set_metrics(m)
    out.update(map=m, noisePower=noise, maxPower=mx)
    if det is not None:
        d1 = cfar_1d(m, g.delay, g.doppler, noise, det["pfa"], det["nGuard"], det["nTrain"], det["minDelay"],
                     det["minDoppler"])
        tcpi = float(g.n) / float(g.fs)
        d2 = centroid(*d1, det["nCentroid"], det["nCentroid"], 1.0 / tcpi)  # blah2.cpp:183
        d3 = interpolate(*d2, m, g.delay, g.doppler, noise, True, True)  # blah2.cpp:178
        out.update(cfar=d1, centroid=d2, detections=d3)
    return out
