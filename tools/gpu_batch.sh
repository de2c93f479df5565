#!/bin/bash
# one GPU call's worth of checks (round 2 working script): the split Toeplitz solve kernel; results under gpurun_out/
O=gpurun_out
mkdir -p $O
timeout 240 python -m pytest tests/test_wh_gpu.py -m gpu -x -q -k "solve or toeplitz" 2>&1 | tail -3 | tee $O/r02s_pytest_solve.log
if ! grep -q " passed" $O/r02s_pytest_solve.log || grep -q "failed\|error" $O/r02s_pytest_solve.log; then echo "solve tests not green: stop"; exit 0; fi
timeout 60 python tools/wh_times.py 2>&1 | tail -1 | tee $O/r02s_wh_times.log
B200DD_WH_SOLVE_SPLIT=0 timeout 60 python tools/wh_times.py 2>&1 | tail -1 | tee -a $O/r02s_wh_times.log
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/r02s_pytest.log
timeout 200 python bench.py --no-cpu-baseline > $O/r02s_bench.json 2> $O/r02s_bench.err; cut -c1-400 $O/r02s_bench.json
timeout 100 python tools/bench_cfg5.py 5 2>/dev/null | tail -1 | tee $O/r02s_cfg5_n1.log
timeout 100 ncu --set full --clock-control none --import-source on -k "regex:wh_solve" --launch-skip 2 -c 1 -f -o $O/r02s_full_solve \
    python tools/wh_times.py 3 > $O/r02s_ncu.log 2>&1; tail -2 $O/r02s_ncu.log
exit 0
