"""Target process for ncu captures (never a source of bench numbers).

  python tools/profile_target.py cfg2 [iters]   one Pipeline (bench.py's workload), `iters` CPIs on one stream
  python tools/profile_target.py cfg3 [iters]   BASELINE configs[2]: CAF only, 2 s CPI @ 10 MS/s, 512 x 1025
  python tools/profile_target.py spectrum [iters]   SpectrumAnalyser(n, 2000) at n = 2e6 and 2e7 (3 kernels each)

Every CPI launches the same kernel sequence, so `ncu --launch-skip` can step over the warm-up CPIs.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from blah2_b200.process import Ambiguity, Pipeline
from blah2_b200.scene import make_scene


def cfg2(iters):
    import bench
    pipe = Pipeline(**bench.GEOM, clutter=bench.CLUTTER, detection=bench.DET, device=0)
    g = pipe.geometry
    sc = make_scene(bench.N, bench.FS, seed=20260923)
    x0 = torch.from_numpy(sc.x.astype(np.complex64)).cuda()
    y0 = torch.from_numpy(sc.y.astype(np.complex64)).cuda()
    dmap = torch.empty((g.n_doppler_bins, g.n_delay_bins), dtype=torch.complex64, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for i in range(iters):
            pipe.submit_device(torch.roll(x0, 977 * i), torch.roll(y0, 977 * i), dmap, stream.cuda_stream)
            r = pipe.fetch(stream.cuda_stream)
    print("cfg2 done", r.get("n_detections") if isinstance(r, dict) else r)


def cfg3(iters):
    geom = (0, 511, -256, 256, 10000000, 20000000, True)
    amb = Ambiguity(*geom)
    g = amb.geometry
    n = geom[5]
    xs = [torch.randn(n, dtype=torch.complex64, device="cuda") for _ in range(2)]
    ys = [torch.randn(n, dtype=torch.complex64, device="cuda") for _ in range(2)]
    out = torch.empty((g.n_doppler_bins, g.n_delay_bins), dtype=torch.complex64, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for i in range(iters):
            amb.process_device(xs[i % 2], ys[i % 2], out, stream.cuda_stream)
    torch.cuda.synchronize()
    print("cfg3 done", g.range_fft_len, g.range_segments, g.range_parts, g.doppler_fft_len)


def spectrum(iters):
    from blah2_b200.process import SpectrumAnalyser
    for n in (2_000_000, 20_000_000):
        sa = SpectrumAnalyser(n, 2000.0)
        xs = [torch.randn(n, dtype=torch.complex64, device="cuda") for _ in range(2)]
        for i in range(iters):
            sa.process_device(xs[i % 2])
        s = sa.fetch()
        print("spectrum done", n, sa.decimation, sa.nSpectrum, float(np.abs(s).max()))
        sa.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    {"cfg2": cfg2, "cfg3": cfg3, "spectrum": spectrum}[which](iters)
