#!/bin/bash
# scratch: A/B of eager vs graph submission in bench.py (3 alternating repeats), device-resident leg only matters
O=gpurun_out; mkdir -p $O
for r in 1 2 3; do
  B200DD_PIPELINE_GRAPH=0 timeout 300 python bench.py --no-cpu-baseline > $O/c4_eager_$r.json 2>> $O/c4.err
  B200DD_PIPELINE_GRAPH=1 timeout 300 python bench.py --no-cpu-baseline > $O/c4_graph_$r.json 2>> $O/c4.err
done
for f in eager_1 graph_1 eager_2 graph_2 eager_3 graph_3; do python - <<PY
import json
try:
    d=json.load(open("$O/c4_$f.json")); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_rspduo_int16"]["value"], d["kernel_ms"])
except Exception as e: print("$f failed", e)
PY
done
tail -3 $O/c4.err
exit 0
