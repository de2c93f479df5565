"""Target for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, at small sizes, results
checked against the oracle so that a sanitizer-clean run is also a correct one.

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from blah2_b200.process import Ambiguity, Pipeline, SpectrumAnalyser, WienerHopf, WienerHopfChunk
from blah2_b200.scene import make_scene, random_iq
from oracle import blah2_oracle as O


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    # range kernel (TMA-staged, several FFT lengths / parts), Doppler kernel; odd pointers exercise the staging edges
    for geom in [(-3, 20, -50, 50, 10000, 4000, False), (2, 40, -30, 30, 10000, 3000, True), (-10, 120, -5000, 5000, 2000000, 60000, True)]:
        x, y = random_iq(geom[5], 3)
        amb = Ambiguity(*geom)
        m = amb.process(x, y)
        ref, _, _ = O.ambiguity_process(x, y, O.ambiguity_geometry(*geom))
        assert rel(m.data, ref) < 1e-5
        n = geom[5]
        bx = torch.empty(n + 1, dtype=torch.complex64, device="cuda")
        by = torch.empty(n + 1, dtype=torch.complex64, device="cuda")
        bx[1:].copy_(torch.from_numpy(x.astype(np.complex64)))
        by[1:].copy_(torch.from_numpy(y.astype(np.complex64)))
        out = torch.zeros((amb.geometry.n_doppler_bins, amb.geometry.n_delay_bins), dtype=torch.complex64, device="cuda")
        torch.cuda.synchronize()
        amb.process_device(bx[1:], by[1:], out)
        torch.cuda.synchronize()
        assert rel(out.cpu().numpy(), ref) < 1e-5
        amb.close()
    # WienerHopf: correlation (TMA staging + wrap fallback), both solve kernels, persistent filter kernel; chunk mode
    for n, dm, dM in [(20011, -10, 40), (60000, -10, 400), (9000, 2, 60)]:
        sc = make_scene(n, 2e6, seed=n % 97, n_clutter=8)
        ok, yf = WienerHopf(dm, dM, n).process(sc.x, sc.y)
        okr, yr = O.wienerhopf_process(sc.x, sc.y, dm, dM)
        assert ok and okr and rel(yf, yr) < 1e-9
        dx = torch.from_numpy(sc.x.astype(np.complex64)).cuda()
        dy = torch.from_numpy(sc.y.astype(np.complex64)).cuda()
        o = torch.empty_like(dy)
        wh = WienerHopf(dm, dM, n)
        torch.cuda.synchronize()
        wh.process_device(dx, dy, o)
        torch.cuda.synchronize()
        assert rel(o.cpu().numpy(), yr) < 1e-5
    n, dm, dM = 60000, -10, 400
    sc = make_scene(n, 2e6, seed=5, n_clutter=8)
    xg, yg = sc.x.astype(np.complex64), sc.y.astype(np.complex64)
    parts, chunks = [], []
    for r in range(2):
        c0, nc = (0, 30000) if r == 0 else (30000, 30000)
        ch = WienerHopfChunk(dm, dM, n, c0, nc)
        xl, xr, yr_ = ch.halos()
        x_loc = torch.from_numpy(xg[np.arange(c0 - xl, c0 + nc + xr) % n]).cuda()
        y_loc = torch.from_numpy(yg[np.arange(c0, c0 + nc + yr_) % n]).cuda()
        ab = torch.zeros(2 * ch.nBins, dtype=torch.complex128, device="cuda")
        torch.cuda.synchronize()
        ch.corr_device(x_loc, y_loc, ab)
        torch.cuda.synchronize()
        parts.append(ab); chunks.append((ch, x_loc, y_loc, nc))
    ab = parts[0] + parts[1]
    outs = []
    for ch, x_loc, y_loc, nc in chunks:
        o = torch.empty(nc, dtype=torch.complex64, device="cuda")
        torch.cuda.synchronize()
        ch.filter_device(ab, x_loc, y_loc, o)
        torch.cuda.synchronize()
        outs.append(o)
    okr, yr = O.wienerhopf_process(xg.astype(np.complex128), yg.astype(np.complex128), dm, dM)
    assert rel(torch.cat(outs).cpu().numpy(), yr) < 1e-5
    # whole chain with detection (metrics, CFAR, scan, emit, fused tail), graph replay, int16 ingest
    fs, n = 200000, 40000
    geom = (-5, 60, -500, 500, fs, n, True)
    det = dict(pfa=1e-4, nGuard=2, nTrain=6, minDelay=3, minDoppler=15.0, nCentroid=4)
    sc = make_scene(n, fs, seed=11)
    pipe = Pipeline(*geom[:6], roundHamming=True, clutter=(-5, 30), detection=det)
    out = pipe.process(sc.x, sc.y)
    ref = O.chain(sc.x, sc.y, O.ambiguity_geometry(*geom), clutter=(-5, 30), det=det)
    assert rel(out["map"], ref["map"]) < 1e-5 and out["detections"].get_nDetections() == len(ref["detections"][0])
    dx = torch.from_numpy(sc.x.astype(np.complex64)).cuda()
    dy = torch.from_numpy(sc.y.astype(np.complex64)).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        pipe.prepare_device(dx, dy, None, s.cuda_stream)
        pipe.submit_device(dx, dy, None, s.cuda_stream)
        r = pipe.fetch(s.cuda_stream)
    assert r["detections"].get_nDetections() == len(ref["detections"][0])
    # spectrum
    x, _ = random_iq(6000, 1)
    sa = SpectrumAnalyser(6000, 2000.0)
    sp = sa.process(x)
    rs, _ = O.spectrum_process(x, 6000, 2000.0)
    assert rel(np.asarray(sp[0] if isinstance(sp, tuple) else sp), rs) < 1e-10
    print("SANITIZE_TARGET OK")


if __name__ == "__main__":
    main()
