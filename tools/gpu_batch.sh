#!/bin/bash
# one GPU call's worth of checks (round 2 working script): the split Toeplitz solve kernel; results under gpurun_out/
O=gpurun_out
T=${1:-r02t}
mkdir -p $O
timeout 200 python -m pytest tests/test_wh_gpu.py -m gpu -x -q -k "solve or toeplitz" 2>&1 | tail -3 | tee $O/${T}_pytest_solve.log
B200DD_WH_SOLVE_SPLIT=1 timeout 60 python tools/wh_times.py 2>&1 | tail -1 | tee $O/${T}_wh_times.log
B200DD_WH_SOLVE_SPLIT=1 timeout 100 ncu --set full --clock-control none --import-source on -k "regex:wh_solve" --launch-skip 2 -c 1 -f -o $O/${T}_full_solve \
    python tools/wh_times.py 3 > $O/${T}_ncu.log 2>&1; tail -1 $O/${T}_ncu.log
B200DD_WH_SOLVE_SPLIT=1 timeout 150 compute-sanitizer --tool memcheck --print-limit 20 tools/sanitize_native > $O/${T}_sanitizer_memcheck_split.log 2>&1; tail -2 $O/${T}_sanitizer_memcheck_split.log
B200DD_WH_SOLVE_SPLIT=1 timeout 150 compute-sanitizer --tool racecheck --print-limit 20 tools/sanitize_native > $O/${T}_sanitizer_racecheck_split.log 2>&1; tail -2 $O/${T}_sanitizer_racecheck_split.log
exit 0
