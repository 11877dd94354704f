#!/bin/bash
O=gpurun_out/final3; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "n2 rc=$?"; cat $O/bench_n2.json | cut -c1-1500; tail -n 3 $O/bench_n2.err
