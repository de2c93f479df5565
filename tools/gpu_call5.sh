#!/bin/bash
# scratch: validate the last spectrum-kernel changes, refresh the bench line and the spectrum capture
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/r01z_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $O/r01z_pytest.log 2>&1
timeout 400 python bench.py > $O/r01z_bench.json 2> $O/r01z_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spec_ -c 12 -f -o $O/r01z_full_spectrum \
    python tools/profile_target.py spectrum 2 > $O/r01z_ncu_spectrum.log 2>&1
cat $O/r01z_pytest.log; cut -c1-300 $O/r01z_bench.json; python -c "
import json; d=json.load(open('$O/r01z_bench.json')); print(d['value'], d['spectrum'])"
exit 0
