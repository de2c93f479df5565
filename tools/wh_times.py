"""Kernel times of the WienerHopf stages at bench.py's workload (CUDA events around each stage, device-resident
float2 IQ): python tools/wh_times.py [iters]   -- a tuning aid, never a source of bench numbers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from blah2_b200.process import WienerHopf
from blah2_b200.scene import make_scene

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sc = make_scene(bench.N, bench.FS, seed=20260923)
x0 = torch.from_numpy(sc.x.astype(np.complex64)).cuda()
y0 = torch.from_numpy(sc.y.astype(np.complex64)).cuda()
xs = [torch.roll(x0, 977 * b) for b in range(10)]
ys = [torch.roll(y0, 977 * b) for b in range(10)]
wh = WienerHopf(bench.CLUTTER[0], bench.CLUTTER[1], bench.N)
yf = torch.empty_like(y0)
s = torch.cuda.Stream()
c, so, ap = [], [], []
with torch.cuda.stream(s):
    for i in range(iters + 3):
        a, b, d = wh.profile_device(xs[i % 10], ys[i % 10], yf, s.cuda_stream)
        if i >= 3:
            c.append(a); so.append(b); ap.append(d)
print("env", {k: v for k, v in os.environ.items() if k.startswith("B200DD")}, "corr %.2f us  solve %.2f us  apply %.2f us" % (
    1e3 * np.mean(c), 1e3 * np.mean(so), 1e3 * np.mean(ap)))
