"""GPU parity: CFAR -> Centroid -> Interpolate and Map::set_metrics vs the oracle.

The reference has no tests for these classes; the oracle's versions are
checked bit-for-bit against the compiled reference (tests/test_oracle_pins.py).
"""
import numpy as np
import pytest

from blah2_b200 import capi
from blah2_b200.process import (Ambiguity, CfarDetector1D, Centroid, Interpolate, Map, WienerHopf, _DetHandle,
                                set_metrics)
from blah2_b200.scene import make_scene, Target
from oracle import blah2_oracle as O

pytestmark = pytest.mark.gpu

DET = dict(pfa=1e-5, nGuard=2, nTrain=6, minDelay=5, minDoppler=15.0, nCentroid=6)


def _oracle_map(seed=5):
    fs, n = 2000000, 200000
    geom = (-10, 120, -5000, 5000, fs, n, True)
    sc = make_scene(n, fs, seed=seed, targets=[Target(37, 3000.0, -30.0), Target(92, -2000.0, -35.0)])
    g = O.ambiguity_geometry(*geom)
    out = O.chain(sc.x, sc.y, g, clutter=(-10, 60), det=DET)
    return geom, sc, g, out


def _same(det, ref, snr_tol=1e-9, pos_tol=0.0):
    assert det.get_nDetections() == len(ref[0]), (det.get_nDetections(), len(ref[0]))
    if len(ref[0]) == 0:
        return
    assert np.max(np.abs(det.delay - ref[0])) <= pos_tol
    assert np.max(np.abs(det.doppler - ref[1])) <= pos_tol
    assert np.max(np.abs(det.snr - ref[2])) <= snr_tol


def test_stages_match_oracle_on_the_same_map():
    geom, sc, g, out = _oracle_map()
    m = Map(out["map"], g.delay, g.doppler, out["noisePower"], out["maxPower"])
    tcpi = geom[5] / geom[4]
    d1 = CfarDetector1D(DET["pfa"], DET["nGuard"], DET["nTrain"], DET["minDelay"], DET["minDoppler"]).process(m)
    assert d1.get_nDetections() > 10
    _same(d1, out["cfar"])                       # positions exact, snr 1e-9 dB
    d2 = Centroid(DET["nCentroid"], DET["nCentroid"], 1.0 / tcpi).process(d1)
    _same(d2, out["centroid"])
    d3 = Interpolate(True, True).process(d2, m)
    _same(d3, out["detections"], snr_tol=1e-9, pos_tol=1e-9)
    assert d3.get_nDetections() >= 2


@pytest.mark.parametrize("flags", [(True, False), (False, True), (False, False)])
def test_interpolate_flag_combinations(flags):
    geom, sc, g, out = _oracle_map()
    m = Map(out["map"], g.delay, g.doppler, out["noisePower"], out["maxPower"])
    ref = O.interpolate(*out["centroid"], out["map"], g.delay, g.doppler, out["noisePower"], *flags)
    from blah2_b200.process import Detection
    d = Interpolate(*flags).process(Detection(*out["centroid"]), m)
    _same(d, ref, snr_tol=1e-9, pos_tol=1e-9)


def test_set_metrics_matches_oracle():
    geom, sc, g, out = _oracle_map()
    m = set_metrics(Map(out["map"], g.delay, g.doppler))
    # map is cast to complex64 for the device copy: 1e-3 dB is the reference's own tolerance
    assert abs(m.noisePower - out["noisePower"]) < 1e-3
    assert abs(m.maxPower - out["maxPower"]) < 1e-3


@pytest.mark.parametrize("params", [
    dict(pfa=1e-12, nGuard=2, nTrain=6, minDelay=5, minDoppler=15.0),       # nothing detected
    dict(pfa=1e-5, nGuard=2, nTrain=6, minDelay=5, minDoppler=1e9),         # every row skipped
    dict(pfa=1e-5, nGuard=2, nTrain=0, minDelay=5, minDoppler=15.0),        # no training cells -> NaN threshold
    dict(pfa=1e-2, nGuard=1, nTrain=3, minDelay=-128, minDoppler=0.0),      # many detections, edges included
    dict(pfa=1e-3, nGuard=258, nTrain=4, minDelay=300, minDoppler=0.0),     # int8_t narrowing of 258 / 300
])
def test_cfar_edge_parameters(params):
    geom, sc, g, out = _oracle_map(seed=6)
    m = Map(out["map"], g.delay, g.doppler, out["noisePower"], out["maxPower"])
    ref = O.cfar_1d(out["map"], g.delay, g.doppler, out["noisePower"], params["pfa"], params["nGuard"],
                    params["nTrain"], params["minDelay"], params["minDoppler"])
    d = CfarDetector1D(params["pfa"], params["nGuard"], params["nTrain"], params["minDelay"],
                       params["minDoppler"]).process(m)
    _same(d, ref)


def test_centroid_uint16_wrap_and_empty_input():
    from blah2_b200.process import Detection
    delay = np.array([2.0, 3.0, 4.0, 40.0, 41.0, 300.0])
    dop = np.array([10.0, 11.0, 10.5, -20.0, -20.5, 0.0])
    snr = np.array([5.0, 9.0, 7.0, 3.0, 4.0, 1.0])
    ref = O.centroid(delay, dop, snr, 6, 6, 1.0)     # delay - 6 < 0 wraps to ~65532 (Centroid.cpp:28)
    d = Centroid(6, 6, 1.0).process(Detection(delay, dop, snr))
    _same(d, ref)
    e = Centroid(6, 6, 1.0).process(Detection(np.zeros(0), np.zeros(0), np.zeros(0)))
    assert e.get_nDetections() == 0


def test_end_to_end_chain_on_device_matches_oracle(relerr):
    """WienerHopf -> Ambiguity -> set_metrics -> CFAR -> Centroid -> Interpolate with the map
    staying on the device between stages (src/blah2.cpp:268-287)."""
    import torch
    geom, sc, g, out = _oracle_map(seed=8)
    fs, n = geom[4], geom[5]
    wh = WienerHopf(-10, 60, n)
    amb = Ambiguity(*geom)
    det = _DetHandle(pfa=DET["pfa"], nGuard=DET["nGuard"], nTrain=DET["nTrain"], minDelay=DET["minDelay"],
                     minDoppler=DET["minDoppler"], nCentroidDelay=DET["nCentroid"], nCentroidDoppler=DET["nCentroid"],
                     resolutionDoppler=1.0 / (n / fs), max_doppler_bins=amb.get_n_doppler_bins(),
                     max_delay_bins=amb.get_n_delay_bins())
    dx = torch.from_numpy(sc.x.astype(np.complex64)).cuda()
    dy = torch.from_numpy(sc.y.astype(np.complex64)).cuda()
    dmap = torch.empty((amb.get_n_doppler_bins(), amb.get_n_delay_bins()), dtype=torch.complex64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        wh.process_device(dx, dy, dy, s.cuda_stream)
        amb.process_device(dx, dy, dmap, s.cuda_stream)
        noise, mx = det.set_metrics_device(dmap, dmap.shape[0], dmap.shape[1], s.cuda_stream)
        d = det.process_device_map(dmap, dmap.shape[0], dmap.shape[1], amb.delay, amb.doppler, noise,
                                   capi.DET_INTERPOLATE, s.cuda_stream)
    e = relerr(dmap.cpu().numpy().astype(np.complex128), out["map"])
    assert e[0] < 1e-5 and e[1] < 1e-5, f"map {e}"
    assert abs(noise - out["noisePower"]) < 1e-3 and abs(mx - out["maxPower"]) < 1e-3
    ref = out["detections"]
    assert d.get_nDetections() == len(ref[0])
    assert np.max(np.abs(d.delay - ref[0])) < 1e-3
    assert np.max(np.abs(d.doppler - ref[1])) < 1e-3 * abs(g.doppler[1] - g.doppler[0]) + 1e-6
    assert np.max(np.abs(d.snr - ref[2])) < 1e-3
