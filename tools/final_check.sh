#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/r01z_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $O/r01z_pytest.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $O/r01z_bench_check.json 2> $O/r01z_bench_check.err
cat $O/r01z_pytest.log; cut -c1-200 $O/r01z_bench_check.json; tail -2 $O/r01z_bench_check.err
exit 0
