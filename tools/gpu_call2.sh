#!/bin/bash
# scratch: second GPU call (spectrum v2 + segment groups in the range kernel)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_spectrum_gpu.py tests/test_caf_gpu.py -x -q 2>&1 | tail -15 > $O/c2_pytest.log
timeout 300 python tools/time_caf.py cfg1 cfg3 cfg4 --groups > $O/c2_time_caf_groups.log 2>&1
for g in 1 2 3; do
  B200DD_CAF_GROUPS=$g timeout 300 python bench.py --no-cpu-baseline > $O/c2_bench_g$g.json 2>> $O/c2_bench.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spec_ -c 12 -f -o $O/c2_full_spectrum \
    python tools/profile_target.py spectrum 2 > $O/c2_ncu_spectrum.log 2>&1
cat $O/c2_pytest.log; cat $O/c2_time_caf_groups.log; for f in g1 g2 g3; do python - <<PY
import json
try:
    d=json.load(open("$O/c2_bench_$f.json")); print("$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["kernel_ms"], d.get("spectrum"))
except Exception as e: print("$f failed", e)
PY
done
tail -3 $O/c2_bench.err; grep -v PROF $O/c2_ncu_spectrum.log | tail -3
exit 0
