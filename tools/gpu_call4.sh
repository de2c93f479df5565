#!/bin/bash
# scratch: A/B of eager vs graph submission in bench.py (3 alternating repeats), device-resident leg only matters
O=gpurun_out; mkdir -p $O
for r in 1 2 3; do
  B200DD_PIPELINE_GRAPH=0 timeout 300 python bench.py --no-cpu-baseline > $O/c4_eager_$r.json 2>> $O/c4.err
  B200DD_PIPELINE_GRAPH=1 timeout 300 python bench.py --no-cpu-baseline > $O/c4_graph_$r.json 2>> $O/c4.err
done
B200DD_PIPELINE_GRAPH=0 B200DD_WH_RADIX=8 timeout 300 python bench.py --no-cpu-baseline > $O/c4_radix8.json 2>> $O/c4.err
for v in 0 1 2; do
  B200DD_SPEC_FOLD=$v timeout 120 python - > $O/c4_fold_$v.log 2>&1 <<PY
import torch, numpy as np, json
from blah2_b200.process import SpectrumAnalyser
for n in (2_000_000, 20_000_000):
    sa = SpectrumAnalyser(n, 2000.0)
    xs = [torch.randn(n, dtype=torch.complex64, device="cuda") for _ in range(3)]
    f, r = [], []
    for i in range(23):
        a, b = sa.profile_device(xs[i % 3])
        if i >= 3: f.append(a); r.append(b)
    print(json.dumps(dict(variant=$v, n=n, fold_us=round(float(np.mean(f))*1e3,2), rest_us=round(float(np.mean(r))*1e3,2), gbs=round(8*sa.nfft/np.mean(f)/1e6,1))))
PY
done
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > $O/c4_pytest_full.log
for f in eager_1 graph_1 eager_2 graph_2 eager_3 graph_3; do python - <<PY
import json
try:
    d=json.load(open("$O/c4_$f.json")); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_rspduo_int16"]["value"], d["kernel_ms"])
except Exception as e: print("$f failed", e)
PY
done
python -c "
import json; d=json.load(open('$O/c4_radix8.json')); print('radix8', d['value'], d['kernel_ms'])"
cat $O/c4_fold_*.log; cat $O/c4_pytest_full.log
tail -3 $O/c4.err
exit 0
