#!/bin/bash
# one GPU call's worth of checks (round 2 working script): the split Toeplitz solve kernel; results under gpurun_out/
O=gpurun_out
T=${1:-r02t}
mkdir -p $O
timeout 200 python -m pytest tests/test_wh_gpu.py -m gpu -x -q -k "solve or toeplitz" 2>&1 | tail -3 | tee $O/${T}_pytest_solve.log
if ! grep -q " passed" $O/${T}_pytest_solve.log || grep -q "failed\|error" $O/${T}_pytest_solve.log; then echo "solve tests not green: stop"; exit 0; fi
timeout 60 python tools/wh_times.py 2>&1 | tail -1 | tee $O/${T}_wh_times.log
B200DD_WH_SOLVE_SPLIT=0 timeout 60 python tools/wh_times.py 2>&1 | tail -1 | tee -a $O/${T}_wh_times.log
timeout 100 ncu --set full --clock-control none --import-source on -k "regex:wh_solve" --launch-skip 2 -c 1 -f -o $O/${T}_full_solve \
    python tools/wh_times.py 3 > $O/${T}_ncu.log 2>&1; tail -1 $O/${T}_ncu.log
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_k20.json 2> $O/${T}_bench.err; cut -c1-200 $O/${T}_bench_k20.json
timeout 120 python bench.py --streams 1 --no-cpu-baseline > $O/${T}_bench_s1.json 2>> $O/${T}_bench.err; cut -c1-200 $O/${T}_bench_s1.json
timeout 100 python tools/bench_cfg5.py 5 2>/dev/null | tail -1 | tee $O/${T}_cfg5_n1.log
exit 0
