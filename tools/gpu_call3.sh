#!/bin/bash
# scratch: third GPU call (CUDA-graph replay of the device chain, spectrum DFT v3)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_spectrum_gpu.py tests/test_pipeline_gpu.py tests/test_caf_gpu.py -x -q -m gpu 2>&1 | tail -15 > $O/c3_pytest.log
timeout 400 python tools/time_e2e.py 36 > $O/c3_time_e2e.json 2> $O/c3_time_e2e.err
B200DD_PIPELINE_GRAPH=0 timeout 300 python bench.py --no-cpu-baseline > $O/c3_bench_eager.json 2>> $O/c3_bench.err
timeout 300 python bench.py --no-cpu-baseline > $O/c3_bench_graph.json 2>> $O/c3_bench.err
timeout 300 python bench.py --no-cpu-baseline --streams 8 > $O/c3_bench_graph_s8.json 2>> $O/c3_bench.err
timeout 200 python tools/time_caf.py cfg1 cfg3 > $O/c3_time_caf.log 2>&1
cat $O/c3_pytest.log; cat $O/c3_time_e2e.json; tail -3 $O/c3_time_e2e.err; for f in eager graph graph_s8; do python - <<PY
import json
try:
    d=json.load(open("$O/c3_bench_$f.json")); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_rspduo_int16"]["value"], d["spectrum"]["n2e6"], d["spectrum"]["n2e7"])
except Exception as e: print("$f failed", e)
PY
done
cat $O/c3_time_caf.log; tail -3 $O/c3_bench.err
exit 0
