#!/bin/bash
# one GPU call's worth of checks (round 2 working script): results under gpurun_out/
O=gpurun_out
tools/ubench/cvt_rates > $O/r02_cvt_rates.log 2>&1; cat $O/r02_cvt_rates.log
python -m pytest tests/test_wh_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/wh_times.py
python tools/sanitize_target.py 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > $O/r02_sanitizer_memcheck.log 2>&1; tail -3 $O/r02_sanitizer_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py > $O/r02_sanitizer_racecheck.log 2>&1; tail -3 $O/r02_sanitizer_racecheck.log
ncu --set full --clock-control none --import-source on -k "regex:caf_range" --launch-skip 2 -c 1 -f -o $O/r02g_full_cfg3 python tools/profile_target.py cfg3 3 > $O/r02g_ncu.log 2>&1; tail -2 $O/r02g_ncu.log
