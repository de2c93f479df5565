#!/bin/bash
# Round-end measurement set, run on the GPU box: tools/round_profile.sh TAG  (outputs under gpurun_out/)
TAG=${1:-r01z}
O=gpurun_out
mkdir -p $O
KF='regex:caf_|wh_|metrics_|cfar_|det_|pl_|scan_|centroid_|compact_|interp_|fft_|spec_'
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $O/${TAG}_pytest.log 2>&1
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python bench.py --streams 1 --no-cpu-baseline > $O/${TAG}_bench_s1.json 2>> $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
# launch list of the bench command (fewer steps: every launch is replayed by ncu; graph replay off so that the list
# holds the steps' own launches and not the plan-creation runs of b200dd_pipeline_prepare_device)
B200DD_PIPELINE_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 1200 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
# full sections: one cfg2 CPI (13 kernels) after two warm-up CPIs; one cfg3 CAF after two warm-up maps
ncu --set full --clock-control none --import-source on -k "$KF" --launch-skip 24 -c 12 -f -o $O/${TAG}_full_cfg2 \
    python tools/profile_target.py cfg2 3 > $O/${TAG}_ncu_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k "$KF" --launch-skip 4 -c 2 -f -o $O/${TAG}_full_cfg3 \
    python tools/profile_target.py cfg3 3 > $O/${TAG}_ncu_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spec_ -c 12 -f -o $O/${TAG}_full_spectrum \
    python tools/profile_target.py spectrum 2 > $O/${TAG}_ncu_spectrum.log 2>&1
python tools/time_caf.py cfg1 cfg3 cfg4 > $O/${TAG}_time_cfg3.log 2>&1
python tools/time_e2e.py 36 > $O/${TAG}_time_e2e.json 2> $O/${TAG}_time_e2e.err
tail -2 $O/${TAG}_pytest.log; cut -c1-400 $O/${TAG}_bench.json; cut -c1-300 $O/${TAG}_bench_reference.json; tail -n 3 $O/${TAG}_ncu_cfg2.log; tail -n 3 $O/${TAG}_ncu_cfg3.log; exit 0
