#!/bin/bash
# two-GPU check of the bench arm (one box): tools/scale2.sh
O=gpurun_out; mkdir -p $O
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $O/r01z_scale_n2.json 2> $O/r01z_scale_n2.err
cut -c1-700 $O/r01z_scale_n2.json; grep -v "^\*\|OMP_NUM" $O/r01z_scale_n2.err | tail -4
exit 0
