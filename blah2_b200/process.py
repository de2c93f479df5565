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
