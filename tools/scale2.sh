#!/bin/bash
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $O/r01z_scale_n2.json 2> $O/r01z_scale_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/r01z_scale_n2_ref.json 2>> $O/r01z_scale_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_cfg5.py > $O/r01z_cfg5_n2.log 2>&1
cut -c1-600 $O/r01z_scale_n2.json; cut -c1-300 $O/r01z_scale_n2_ref.json; tail -3 $O/r01z_cfg5_n2.log; tail -5 $O/r01z_scale_n2.err
exit 0
