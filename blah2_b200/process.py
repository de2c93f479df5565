"""Host-side mirror of the reference's process classes over the C ABI.

Same class names, constructor arguments, getters and error behaviour as
``src/process/{ambiguity,clutter,detection}`` in the reference, so the parity tests read
like the reference's own unit tests (test/unit/process/ambiguity/TestAmbiguity.cpp).
Arrays replace the reference's containers: ``IqData`` -> 1-D complex arrays,
``Map<complex<double>>`` -> :class:`Map`, ``Detection`` -> :class:`Detection`.

Every ``process`` call runs CUDA kernels through libb200dd.so.  There is no CPU
fallback; missing library / device raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass
class Map:
    """Mirror of Map<std::complex<double>> (src/data/Map.h:30-42)."""
    data: np.ndarray  # [nRows = nDopplerBins][nCols = nDelayBins] complex128
    delay: np.ndarray  # int32 bins
    doppler: np.ndarray  # float64 Hz
    noisePower: float = 0.0
    maxPower: float = 0.0

    def get_nRows(self):
        return self.data.shape[0]

    def get_nCols(self):
        return self.data.shape[1]


@dataclass
class Detection:
    """Mirror of Detection (src/data/Detection.h:17-24)."""
    delay: np.ndarray
    doppler: np.ndarray
    snr: np.ndarray

    def get_nDetections(self):
        return int(self.delay.shape[0])


def _c128(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.complex128))


class Ambiguity:
    """Ambiguity(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming=False)
    -- src/process/ambiguity/Ambiguity.h:34."""

    def __init__(self, delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming=False, device=-1):
        lib = capi.load()
        p = capi.CafParams(int(delayMin), int(delayMax), int(dopplerMin), int(dopplerMax), int(fs), int(n),
                           int(bool(roundHamming)), int(device))
        h = C.c_void_p()
        capi.check(lib.b200dd_caf_create(C.byref(p), C.byref(h)))
        self._lib, self._h = lib, h
        g = capi.CafGeometry()
        capi.check(lib.b200dd_caf_get_geometry(h, C.byref(g)))
        self.geometry = g
        self.delay = np.empty(g.n_delay_bins, dtype=np.int32)
        self.doppler = np.empty(g.n_doppler_bins, dtype=np.float64)
        capi.check(lib.b200dd_caf_get_axes(h, capi.ptr(self.delay), capi.ptr(self.doppler)))
        self._n_samples = int(n)

    # getters, Ambiguity.h:46-58
    def get_doppler_middle(self):
        return self.geometry.doppler_middle

    def get_n_delay_bins(self):
        return self.geometry.n_delay_bins

    def get_n_doppler_bins(self):
        return self.geometry.n_doppler_bins

    def get_n_corr(self):
        return self.geometry.n_corr

    def get_cpi(self):
        return self.geometry.cpi

    def get_nfft(self):
        return self.geometry.nfft

    def get_n_samples(self):
        return self._n_samples

    @property
    def n_used(self):
        return self.geometry.n_used

    def process(self, x, y) -> Map:
        """Host path: x, y complex arrays (>= n_used samples).  Returns the Map; like the
        reference only the first nDopplerBins*nCorr samples are consumed and
        get_n_samples() afterwards reports that count (Ambiguity.cpp:105)."""
        x, y = _c128(x), _c128(y)
        if x.shape[0] != y.shape[0]:
            raise ValueError("x and y must have the same length")
        g = self.geometry
        out = np.empty((g.n_doppler_bins, g.n_delay_bins), dtype=np.complex128)
        capi.check(self._lib.b200dd_caf_process_host(self._h, capi.ptr(x), capi.ptr(y), x.shape[0], capi.ptr(out)))
        self._n_samples = g.n_used
        return Map(out, self.delay.copy(), self.doppler.copy())

    def process_device(self, d_x, d_y, d_map=None, stream=None):
        """Device path: torch complex64 CUDA tensors (or raw device pointers as ints).
        Asynchronous.  Returns None; result in d_map (or the handle's internal map)."""
        n = d_x.numel() if hasattr(d_x, "numel") else self.geometry.n_used
        capi.check(self._lib.b200dd_caf_process_device(self._h, capi.ptr(d_x), capi.ptr(d_y), int(n),
                                                       capi.ptr(d_map), capi.ptr(stream) if stream else None))

    def range_device(self, d_x, d_y, batch0, n_batches, d_R, stream=None):
        """Range stage on batches [batch0, batch0+n_batches): d_x, d_y hold exactly those batches."""
        capi.check(self._lib.b200dd_caf_range_device(self._h, capi.ptr(d_x), capi.ptr(d_y), int(batch0), int(n_batches),
                                                     capi.ptr(d_R), capi.ptr(stream) if stream else None))

    def doppler_device(self, d_R, col0, n_cols, d_map_tile, stream=None):
        """Doppler stage on delay columns [col0, col0+n_cols) of the complete range matrix d_R."""
        capi.check(self._lib.b200dd_caf_doppler_device(self._h, capi.ptr(d_R), int(col0), int(n_cols),
                                                       capi.ptr(d_map_tile), capi.ptr(stream) if stream else None))

    def place_tile(self, d_tile, col0, n_cols, d_map, stream=None):
        """Tile [nDop][n_cols] -> delay columns [col0, col0 + n_cols) of the row-major map d_map."""
        capi.check(self._lib.b200dd_caf_place_tile_device(self._h, capi.ptr(d_tile), int(col0), int(n_cols), capi.ptr(d_map),
                                                          capi.ptr(stream) if stream else None))

    def place_tiles(self, d_tiles, n_tiles, d_map, stream=None):
        """All gathered tiles of an equal column split (shard.block_range) -> the row-major map, one kernel."""
        capi.check(self._lib.b200dd_caf_place_tiles_device(self._h, capi.ptr(d_tiles), int(n_tiles), capi.ptr(d_map),
                                                           capi.ptr(stream) if stream else None))

    def profile_device(self, d_x, d_y, d_map=None, stream=None):
        """(ms_range, ms_doppler): CUDA-event durations of the two CAF kernels for one CPI."""
        a, b = C.c_float(), C.c_float()
        n = d_x.numel() if hasattr(d_x, "numel") else self.geometry.n_used
        capi.check(self._lib.b200dd_caf_profile_device(self._h, capi.ptr(d_x), capi.ptr(d_y), int(n), capi.ptr(d_map),
                                                       capi.ptr(stream) if stream else None, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_range_matrix(self):
        g = self.geometry
        out = np.empty((g.n_doppler_bins, g.n_delay_bins), dtype=np.complex64)
        capi.check(self._lib.b200dd_caf_debug_range_matrix(self._h, capi.ptr(out)))
        return out

    def device_map_ptr(self) -> int:
        return int(self._lib.b200dd_caf_device_map(self._h) or 0)

    def stream_ptr(self) -> int:
        return int(self._lib.b200dd_caf_stream(self._h) or 0)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200dd_caf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WienerHopf:
    """WienerHopf(delayMin, delayMax, nSamples) -- src/process/clutter/WienerHopf.h:68."""

    def __init__(self, delayMin, delayMax, nSamples, device=-1):
        lib = capi.load()
        h = C.c_void_p()
        capi.check(lib.b200dd_wh_create(int(delayMin), int(delayMax), int(nSamples), int(device), C.byref(h)))
        self._lib, self._h = lib, h
        self.nSamples = int(nSamples)
        self.nBins = int(lib.b200dd_wh_n_bins(h))
        self.plan = capi.WhPlan()
        capi.check(lib.b200dd_wh_get_plan(h, C.byref(self.plan)))

    def process(self, x, y):
        """Host path.  Returns (ok, y_filtered): ok False <=> the reference returns false
        (Cholesky failure) and leaves y untouched (WienerHopf.cpp:111-122)."""
        x = _c128(x)
        y = _c128(y).copy()
        if x.shape[0] != self.nSamples or y.shape[0] != self.nSamples:
            raise ValueError("x and y must hold nSamples samples")
        rc = capi.check(self._lib.b200dd_wh_process_host(self._h, capi.ptr(x), capi.ptr(y)),
                        allow=(capi.FILTER_FAILED,))
        return rc == capi.OK, y

    def process_device(self, d_x, d_y, d_y_out=None, stream=None):
        out = d_y if d_y_out is None else d_y_out
        capi.check(self._lib.b200dd_wh_process_device(self._h, capi.ptr(d_x), capi.ptr(d_y), capi.ptr(out),
                                                      capi.ptr(stream) if stream else None))

    def profile_device(self, d_x, d_y, d_y_out, stream=None):
        """(ms_corr, ms_solve, ms_apply): CUDA-event durations of the filter's stages."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        capi.check(self._lib.b200dd_wh_profile_device(self._h, capi.ptr(d_x), capi.ptr(d_y), capi.ptr(d_y_out),
                                                      capi.ptr(stream) if stream else None, C.byref(a), C.byref(b),
                                                      C.byref(c)))
        return a.value, b.value, c.value

    def last_status(self) -> bool:
        rc = capi.check(self._lib.b200dd_wh_last_status(self._h), allow=(capi.FILTER_FAILED,))
        return rc == capi.OK

    def debug_weights(self):
        w = np.empty(self.nBins, dtype=np.complex128)
        a = np.empty(self.nBins, dtype=np.complex128)
        b = np.empty(self.nBins, dtype=np.complex128)
        capi.check(self._lib.b200dd_wh_debug_weights(self._h, capi.ptr(w), capi.ptr(a), capi.ptr(b)))
        return w, a, b

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200dd_wh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WienerHopfChunk:
    """The clutter filter on ONE CHUNK of a CPI that is split over several GPUs (C ABI
    b200dd_wh_create_chunk).  Device path only (complex64 CUDA tensors).  x_loc / y_loc are the chunk plus the
    halos `halos()` reports; blah2_b200.shard.wienerhopf_single_cpi_sharded drives it across ranks."""

    def __init__(self, delayMin, delayMax, nSamples, chunk_begin, chunk_len, device=-1):
        lib = capi.load()
        h = C.c_void_p()
        capi.check(lib.b200dd_wh_create_chunk(int(delayMin), int(delayMax), int(nSamples), int(chunk_begin), int(chunk_len),
                                              int(device), C.byref(h)))
        self._lib, self._h = lib, h
        self.nSamples, self.chunk_begin, self.chunk_len = int(nSamples), int(chunk_begin), int(chunk_len)
        self.nBins = int(lib.b200dd_wh_n_bins(h))
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        capi.check(lib.b200dd_wh_chunk_halos(h, C.byref(a), C.byref(b), C.byref(c)))
        self.x_left, self.x_right, self.y_right = int(a.value), int(b.value), int(c.value)

    def halos(self):
        """(x_left, x_right, y_right) in samples."""
        return self.x_left, self.x_right, self.y_right

    def corr_device(self, d_x_loc, d_y_loc, d_ab, stream=None):
        """d_ab (complex128 CUDA tensor, 2 nBins) <- this chunk's share of (a, b)."""
        assert d_x_loc.numel() == self.x_left + self.chunk_len + self.x_right and d_y_loc.numel() == self.chunk_len + self.y_right
        capi.check(self._lib.b200dd_wh_chunk_corr_device(self._h, capi.ptr(d_x_loc), capi.ptr(d_y_loc), capi.ptr(d_ab),
                                                         capi.ptr(stream) if stream else None))

    def filter_device(self, d_ab, d_x_loc, d_y_loc, d_y_out, stream=None):
        """Replicated solve on the summed d_ab, then the chunk's chunk_len filtered samples into d_y_out."""
        assert d_y_out.numel() == self.chunk_len
        capi.check(self._lib.b200dd_wh_chunk_filter_device(self._h, capi.ptr(d_ab), capi.ptr(d_x_loc), capi.ptr(d_y_loc),
                                                           capi.ptr(d_y_out), capi.ptr(stream) if stream else None))

    def last_status(self) -> bool:
        rc = capi.check(self._lib.b200dd_wh_last_status(self._h), allow=(capi.FILTER_FAILED,))
        return rc == capi.OK

    def debug_weights(self):
        """(w, a, b) of the last filter_device call (tests)."""
        w = np.empty(self.nBins, dtype=np.complex128)
        a = np.empty(self.nBins, dtype=np.complex128)
        b = np.empty(self.nBins, dtype=np.complex128)
        capi.check(self._lib.b200dd_wh_debug_weights(self._h, capi.ptr(w), capi.ptr(a), capi.ptr(b)))
        return w, a, b

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200dd_wh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpectrumAnalyser:
    """SpectrumAnalyser(n, bandwidth) -- src/process/spectrum/SpectrumAnalyser.h:48.

    ``process(x)`` returns ``(spectrum, frequency)``: the two vectors the reference hands to
    ``IqData::update_spectrum`` / ``update_frequency`` (SpectrumAnalyser.cpp:55,68)."""

    def __init__(self, n, bandwidth, device=-1):
        lib = capi.load()
        h = C.c_void_p()
        capi.check(lib.b200dd_spectrum_create(int(n), float(bandwidth), int(device), C.byref(h)))
        self._lib, self._h = lib, h
        g = capi.SpectrumGeometry()
        capi.check(lib.b200dd_spectrum_get_geometry(h, C.byref(g)))
        self.geometry = g
        self.decimation, self.nSpectrum, self.nfft = int(g.decimation), int(g.n_spectrum), int(g.nfft)
        self.frequency = np.empty(int(g.n_frequency), dtype=np.float64)
        capi.check(lib.b200dd_spectrum_get_frequency(h, capi.ptr(self.frequency), int(g.n_frequency)))

    def process(self, x):
        """Host path: complex array with >= nfft samples (read, not consumed)."""
        x = _c128(x)
        out = np.empty(self.nSpectrum, dtype=np.complex128)
        capi.check(self._lib.b200dd_spectrum_process_host(self._h, capi.ptr(x), x.shape[0], capi.ptr(out)))
        return out, self.frequency.copy()

    def process_device(self, d_x, d_spectrum=None, stream=None):
        """Device path: torch complex64 (or complex128) CUDA tensor.  Asynchronous; pair with fetch()
        when d_spectrum is None."""
        n = d_x.numel()
        f64 = getattr(d_x, "element_size", lambda: 8)() == 16
        fn = self._lib.b200dd_spectrum_process_device_f64 if f64 else self._lib.b200dd_spectrum_process_device
        capi.check(fn(self._h, capi.ptr(d_x), int(n), capi.ptr(d_spectrum), capi.ptr(stream) if stream else None))

    def fetch(self, stream=None):
        out = np.empty(self.nSpectrum, dtype=np.complex128)
        capi.check(self._lib.b200dd_spectrum_fetch(self._h, capi.ptr(out), capi.ptr(stream) if stream else None))
        return out

    def profile_device(self, d_x, stream=None):
        """(ms_fold, ms_rest): CUDA-event durations of the folding pass and of reduction + DFT."""
        a, b = C.c_float(), C.c_float()
        capi.check(self._lib.b200dd_spectrum_profile_device(self._h, capi.ptr(d_x), int(d_x.numel()),
                                                            capi.ptr(stream) if stream else None, C.byref(a),
                                                            C.byref(b)))
        return a.value, b.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200dd_spectrum_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DetHandle:
    """One b200dd_det handle shared by the three detection classes of a pipeline."""

    def __init__(self, pfa=1e-5, nGuard=0, nTrain=0, minDelay=0, minDoppler=0.0, nCentroidDelay=0,
                 nCentroidDoppler=0, resolutionDoppler=1.0, doDelay=True, doDoppler=True, max_doppler_bins=8192,
                 max_delay_bins=2048, device=-1):
        lib = capi.load()
        p = capi.DetParams(float(pfa), int(nGuard), int(nTrain), int(minDelay), float(minDoppler),
                           int(nCentroidDelay), int(nCentroidDoppler), float(resolutionDoppler), int(bool(doDelay)),
                           int(bool(doDoppler)), int(device))
        h = C.c_void_p()
        capi.check(lib.b200dd_det_create(C.byref(p), int(max_doppler_bins), int(max_delay_bins), C.byref(h)))
        self._lib, self._h = lib, h
        self.cap = min(int(max_doppler_bins) * int(max_delay_bins), 1 << 18)

    def _bufs(self, cap):
        return [np.empty(max(1, cap), dtype=np.float64) for _ in range(3)]

    def process_map(self, m: Map, last_stage: int) -> Detection:
        data = _c128(m.data)
        delay = np.ascontiguousarray(m.delay, dtype=np.int32)
        doppler = np.ascontiguousarray(m.doppler, dtype=np.float64)
        od, of, os_ = self._bufs(self.cap)
        n = C.c_uint32()
        capi.check(self._lib.b200dd_det_process_host(self._h, last_stage, capi.ptr(data), data.shape[0],
                                                     data.shape[1], capi.ptr(delay), capi.ptr(doppler),
                                                     float(m.noisePower), capi.ptr(od), capi.ptr(of), capi.ptr(os_),
                                                     self.cap, C.byref(n)))
        k = n.value
        return Detection(od[:k].copy(), of[:k].copy(), os_[:k].copy())

    def process_device_map(self, d_map, nDop, nDel, delay, doppler, noisePower, last_stage, stream=None) -> Detection:
        delay = np.ascontiguousarray(delay, dtype=np.int32)
        doppler = np.ascontiguousarray(doppler, dtype=np.float64)
        od, of, os_ = self._bufs(self.cap)
        n = C.c_uint32()
        capi.check(self._lib.b200dd_det_process_device(self._h, last_stage, capi.ptr(d_map), int(nDop), int(nDel),
                                                       capi.ptr(delay), capi.ptr(doppler), float(noisePower),
                                                       capi.ptr(od), capi.ptr(of), capi.ptr(os_), self.cap,
                                                       C.byref(n), capi.ptr(stream) if stream else None))
        k = n.value
        return Detection(od[:k].copy(), of[:k].copy(), os_[:k].copy())

    def set_metrics_device(self, d_map, nDop, nDel, stream=None):
        out = np.empty(2, dtype=np.float64)
        capi.check(self._lib.b200dd_det_set_metrics_device(self._h, capi.ptr(d_map), int(nDop), int(nDel),
                                                           capi.ptr(out), capi.ptr(stream) if stream else None))
        return float(out[0]), float(out[1])

    def centroid(self, det: Detection) -> Detection:
        d = np.ascontiguousarray(det.delay, dtype=np.float64)
        f = np.ascontiguousarray(det.doppler, dtype=np.float64)
        s = np.ascontiguousarray(det.snr, dtype=np.float64)
        cap = max(1, d.shape[0])
        od, of, os_ = self._bufs(cap)
        n = C.c_uint32()
        capi.check(self._lib.b200dd_det_centroid_host(self._h, capi.ptr(d), capi.ptr(f), capi.ptr(s), d.shape[0],
                                                      capi.ptr(od), capi.ptr(of), capi.ptr(os_), cap, C.byref(n)))
        k = n.value
        return Detection(od[:k].copy(), of[:k].copy(), os_[:k].copy())

    def interpolate(self, det: Detection, m: Map) -> Detection:
        d = np.ascontiguousarray(det.delay, dtype=np.float64)
        f = np.ascontiguousarray(det.doppler, dtype=np.float64)
        s = np.ascontiguousarray(det.snr, dtype=np.float64)
        data = _c128(m.data)
        delay = np.ascontiguousarray(m.delay, dtype=np.int32)
        doppler = np.ascontiguousarray(m.doppler, dtype=np.float64)
        cap = max(1, d.shape[0])
        od, of, os_ = self._bufs(cap)
        n = C.c_uint32()
        capi.check(self._lib.b200dd_det_interpolate_host(self._h, capi.ptr(d), capi.ptr(f), capi.ptr(s), d.shape[0],
                                                         capi.ptr(data), data.shape[0], data.shape[1],
                                                         capi.ptr(delay), capi.ptr(doppler), float(m.noisePower),
                                                         capi.ptr(od), capi.ptr(of), capi.ptr(os_), cap, C.byref(n)))
        k = n.value
        return Detection(od[:k].copy(), of[:k].copy(), os_[:k].copy())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200dd_det_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CfarDetector1D:
    """CfarDetector1D(pfa, nGuard, nTrain, minDelay, minDoppler) -- CfarDetector1D.h:46."""

    def __init__(self, pfa, nGuard, nTrain, minDelay, minDoppler, max_doppler_bins=8192, max_delay_bins=2048,
                 device=-1):
        self._d = _DetHandle(pfa=pfa, nGuard=nGuard, nTrain=nTrain, minDelay=minDelay, minDoppler=minDoppler,
                             max_doppler_bins=max_doppler_bins, max_delay_bins=max_delay_bins, device=device)

    def process(self, m: Map) -> Detection:
        return self._d.process_map(m, capi.DET_CFAR)


class Centroid:
    """Centroid(nDelay, nDoppler, resolutionDoppler) -- Centroid.h:35."""

    def __init__(self, nDelay, nDoppler, resolutionDoppler, device=-1):
        self._d = _DetHandle(nCentroidDelay=nDelay, nCentroidDoppler=nDoppler, resolutionDoppler=resolutionDoppler,
                             max_doppler_bins=512, max_delay_bins=512, device=device)

    def process(self, det: Detection) -> Detection:
        return self._d.centroid(det)


class Interpolate:
    """Interpolate(doDelay, doDoppler) -- Interpolate.h:36."""

    def __init__(self, doDelay, doDoppler, max_doppler_bins=8192, max_delay_bins=2048, device=-1):
        self._d = _DetHandle(doDelay=doDelay, doDoppler=doDoppler, max_doppler_bins=max_doppler_bins,
                             max_delay_bins=max_delay_bins, device=device)

    def process(self, det: Detection, m: Map) -> Detection:
        return self._d.interpolate(det, m)


def set_metrics(m: Map) -> Map:
    """Map::set_metrics (Map.cpp:188-206) evaluated on the GPU for a host Map."""
    import torch  # device memory plumbing only

    d = _DetHandle(max_doppler_bins=m.data.shape[0], max_delay_bins=m.data.shape[1])
    t = torch.from_numpy(np.ascontiguousarray(m.data.astype(np.complex64))).cuda()
    m.noisePower, m.maxPower = d.set_metrics_device(t, m.data.shape[0], m.data.shape[1])
    torch.cuda.synchronize()
    d.close()
    return m


class Pipeline:
    """One CPI through the body of the reference's process thread (src/blah2.cpp:268-287):
    [WienerHopf] -> Ambiguity -> Map::set_metrics -> [CfarDetector1D -> Centroid -> Interpolate],
    intermediates resident on the device.  Parameters are named like config/config.yml."""

    def __init__(self, delayMin, delayMax, dopplerMin, dopplerMax, fs, nSamples, roundHamming=True, clutter=None,
                 detection=None, device=-1, max_detections=4096, spectrum_bandwidth=None):
        lib = capi.load()
        p = capi.PipelineParams()
        p.caf = capi.CafParams(int(delayMin), int(delayMax), int(dopplerMin), int(dopplerMax), int(fs),
                               int(nSamples), int(bool(roundHamming)), int(device))
        p.clutter_enable = int(clutter is not None)
        if clutter is not None:
            p.clutter_delay_min, p.clutter_delay_max = int(clutter[0]), int(clutter[1])
        p.detection_enable = int(detection is not None)
        d = detection or {}
        tcpi = float(nSamples) / float(fs)
        p.det = capi.DetParams(float(d.get("pfa", 1e-5)), int(d.get("nGuard", 2)), int(d.get("nTrain", 6)),
                               int(d.get("minDelay", 5)), float(d.get("minDoppler", 15.0)),
                               int(d.get("nCentroid", 6)), int(d.get("nCentroid", 6)), 1.0 / tcpi, 1, 1, int(device))
        h = C.c_void_p()
        capi.check(lib.b200dd_pipeline_create(C.byref(p), C.byref(h)))
        self._lib, self._h = lib, h
        g = capi.CafGeometry()
        capi.check(lib.b200dd_pipeline_get_geometry(h, C.byref(g)))
        self.geometry = g
        self.delay = np.empty(g.n_delay_bins, dtype=np.int32)
        self.doppler = np.empty(g.n_doppler_bins, dtype=np.float64)
        capi.check(lib.b200dd_pipeline_get_axes(h, capi.ptr(self.delay), capi.ptr(self.doppler)))
        self.cap = int(max_detections)
        self._od, self._of, self._os = (np.empty(self.cap, dtype=np.float64) for _ in range(3))
        self.n_samples = int(nSamples)
        self.n_spectrum = 0
        if spectrum_bandwidth is not None:  # spectrumAnalyser->process(x), blah2.cpp:263-265
            ns = C.c_uint32()
            capi.check(lib.b200dd_pipeline_enable_spectrum(h, float(spectrum_bandwidth), C.byref(ns)))
            self.n_spectrum = int(ns.value)

    def fetch_spectrum(self):
        """Spectrum of the reference channel of the last submitted CPI (complex128, nSpectrum bins)."""
        out = np.empty(self.n_spectrum, dtype=np.complex128)
        capi.check(self._lib.b200dd_pipeline_fetch_spectrum(self._h, capi.ptr(out), self.n_spectrum))
        return out

    def _result(self, res, want_map_arr=None):
        skipped = res.filter_status != capi.OK
        # a failed clutter-filter solve skips the CPI (blah2.cpp:270-273): no detections, no metrics, no map product
        k = 0 if skipped else min(res.n_detections, self.cap)
        det = Detection(self._od[:k].copy(), self._of[:k].copy(), self._os[:k].copy())
        return dict(skipped=skipped, noisePower=res.noise_power, maxPower=res.max_power,
                    detections=det, map=None if skipped else want_map_arr)

    def process(self, x, y, want_map=True, map_out=None):
        """Host path: complex128 arrays (pinned torch tensors / numpy).  Synchronous."""
        g = self.geometry
        if map_out is None and want_map:
            map_out = np.empty((g.n_doppler_bins, g.n_delay_bins), dtype=np.complex128)
        n = x.numel() if hasattr(x, "numel") else x.shape[0]
        res = capi.CpiResult()
        capi.check(self._lib.b200dd_pipeline_process_host(self._h, capi.ptr(x), capi.ptr(y), int(n),
                                                          capi.ptr(map_out) if map_out is not None else None,
                                                          C.byref(res), capi.ptr(self._od), capi.ptr(self._of),
                                                          capi.ptr(self._os), self.cap), allow=(capi.ERR_CAPACITY,))
        return self._result(res, map_out)

    def submit_host(self, x, y, map_out=None):
        """Asynchronous host path: enqueue H2D + kernels + D2H(map) and return; pair with fetch().
        x, y, map_out must stay alive (pinned for real overlap) until fetch() returns."""
        n = x.numel() if hasattr(x, "numel") else x.shape[0]
        self._pending_map = map_out
        capi.check(self._lib.b200dd_pipeline_submit_host(self._h, capi.ptr(x), capi.ptr(y), int(n),
                                                         capi.ptr(map_out) if map_out is not None else None))

    def submit_host_rspduo(self, iq, map_out=None):
        """Asynchronous host path fed with the reference's replay layout: int16 [n, 4] = I1 Q1 I2 Q2
        (blah2_b200/scene.py write_rspduo).  Pair with fetch()."""
        n = (iq.numel() if hasattr(iq, "numel") else iq.size) // 4
        self._pending_map = map_out
        capi.check(self._lib.b200dd_pipeline_submit_host_rspduo(self._h, capi.ptr(iq), int(n),
                                                                capi.ptr(map_out) if map_out is not None else None))

    def submit_device(self, d_x, d_y, d_map=None, stream=None):
        n = d_x.numel() if hasattr(d_x, "numel") else self.n_samples
        capi.check(self._lib.b200dd_pipeline_submit_device(self._h, capi.ptr(d_x), capi.ptr(d_y), int(n),
                                                           capi.ptr(d_map), capi.ptr(stream) if stream else None))

    def prepare_device(self, d_x, d_y, d_map=None, stream=None):
        """Plan creation for submit_device on this buffer triple (records the CUDA graph of the chain now)."""
        n = d_x.numel() if hasattr(d_x, "numel") else self.n_samples
        capi.check(self._lib.b200dd_pipeline_prepare_device(self._h, capi.ptr(d_x), capi.ptr(d_y), int(n),
                                                            capi.ptr(d_map), capi.ptr(stream) if stream else None))

    def fetch(self, stream=None):
        res = capi.CpiResult()
        capi.check(self._lib.b200dd_pipeline_fetch(self._h, C.byref(res), capi.ptr(self._od), capi.ptr(self._of),
                                                   capi.ptr(self._os), self.cap,
                                                   capi.ptr(stream) if stream else None), allow=(capi.ERR_CAPACITY,))
        return self._result(res)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200dd_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
