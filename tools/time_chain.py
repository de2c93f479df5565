"""Scratch timing of the device-resident WienerHopf -> CAF -> detection chain (BASELINE config 2)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from blah2_b200 import capi
from blah2_b200.process import Ambiguity, WienerHopf, _DetHandle
from blah2_b200.scene import make_scene

def main(iters=20, warm=3, nbuf=6):
    fs, n = 2000000, 2000000
    geom = (0, 299, -128, 128, fs, n, True)
    amb = Ambiguity(*geom)
    wh = WienerHopf(-10, 400, n)
    det = _DetHandle(pfa=1e-5, nGuard=2, nTrain=6, minDelay=5, minDoppler=15.0, nCentroidDelay=6, nCentroidDoppler=6,
                     resolutionDoppler=1.0 / (n / fs), max_doppler_bins=amb.get_n_doppler_bins(),
                     max_delay_bins=amb.get_n_delay_bins())
    sc = make_scene(n, fs, seed=1)
    xs, ys = [], []
    for b in range(nbuf):
        xs.append(torch.from_numpy(np.roll(sc.x, 1000 * b).astype(np.complex64)).cuda())
        ys.append(torch.from_numpy(np.roll(sc.y, 1000 * b).astype(np.complex64)).cuda())
    yf = torch.empty_like(ys[0])
    dmap = torch.empty((amb.get_n_doppler_bins(), amb.get_n_delay_bins()), dtype=torch.complex64, device="cuda")
    stream = torch.cuda.Stream()
    st = stream.cuda_stream
    stages = {"wh": [], "caf": [], "det": []}
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for it in range(warm + iters):
            b = it % nbuf
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(stream)
            wh.process_device(xs[b], ys[b], yf, st)
            ev[1].record(stream)
            amb.process_device(xs[b], yf, dmap, st)
            ev[2].record(stream)
            noise, mx = det.set_metrics_device(dmap, dmap.shape[0], dmap.shape[1], st)
            d = det.process_device_map(dmap, dmap.shape[0], dmap.shape[1], amb.delay, amb.doppler, noise,
                                       capi.DET_INTERPOLATE, st)
            ev[3].record(stream)
            stream.synchronize()
            if it >= warm:
                stages["wh"].append(ev[0].elapsed_time(ev[1]))
                stages["caf"].append(ev[1].elapsed_time(ev[2]))
                stages["det"].append(ev[2].elapsed_time(ev[3]))
    out = {k: round(float(np.median(v)), 4) for k, v in stages.items()}
    out["n_det"] = d.get_nDetections()
    out["ok"] = wh.last_status()
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
