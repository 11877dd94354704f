#!/bin/bash
# node-order sensitivity of the large-pair kernel on C4 and what the native reorderings recover
mkdir -p gpurun_out
for o in natural random random+rcm random+pbr; do
  timeout 600 python tools/bench_configs.py --only C4 --c4-graphs 200 --c4-order $o 2>&1 | grep -v "arn" | tail -n 3
done | tee gpurun_out/c4_order.jsonl
