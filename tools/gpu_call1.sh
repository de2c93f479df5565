#!/bin/bash
# scratch: first GPU call of the session (spectrum stage parity + knob experiments + e2e diagnosis)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_spectrum_gpu.py -x -q 2>&1 | tail -15 > $O/c1_pytest_spectrum.log
timeout 300 python tools/time_e2e.py 24 > $O/c1_time_e2e.json 2> $O/c1_time_e2e.err
timeout 300 python bench.py --no-cpu-baseline > $O/c1_bench_default.json 2> $O/c1_bench.err
B200DD_WH_CORR_LOG2M=11 timeout 300 python bench.py --no-cpu-baseline > $O/c1_bench_corr11.json 2>> $O/c1_bench.err
B200DD_WH_APPLY_LOG2M=11 timeout 300 python bench.py --no-cpu-baseline > $O/c1_bench_apply11.json 2>> $O/c1_bench.err
timeout 300 python bench.py --no-cpu-baseline --streams 6 > $O/c1_bench_s6.json 2>> $O/c1_bench.err
timeout 300 python bench.py --no-cpu-baseline --streams 8 > $O/c1_bench_s8.json 2>> $O/c1_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spec_ -c 12 -f -o $O/c1_full_spectrum \
    python tools/profile_target.py spectrum 2 > $O/c1_ncu_spectrum.log 2>&1
cat $O/c1_pytest_spectrum.log; cat $O/c1_time_e2e.json; for f in default corr11 apply11 s6 s8; do python - <<PY
import json
try:
    d=json.load(open("$O/c1_bench_$f.json")); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_rspduo_int16"]["value"], d["kernel_ms"], d.get("spectrum"))
except Exception as e: print("$f failed", e)
PY
done
tail -3 $O/c1_bench.err; tail -3 $O/c1_ncu_spectrum.log
exit 0
