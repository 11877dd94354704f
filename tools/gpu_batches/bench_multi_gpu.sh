#!/bin/bash
# multi-GPU bench: C5 strong scaling at N = $1
N=${1:-2}
mkdir -p gpurun_out/c
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/c/bench_n$N.json 2> gpurun_out/c/bench_n$N.err
echo "rc=$?"
tail -c 3000 gpurun_out/c/bench_n$N.json
tail -n 5 gpurun_out/c/bench_n$N.err
