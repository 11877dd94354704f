#!/bin/bash
# C4 final numbers: full 500-graph set, gradient, node-order sensitivity, ncu
mkdir -p gpurun_out/g
timeout 120 python tools/bench_configs.py --only C4 --c4-graphs 60 2>/dev/null | tee gpurun_out/g/c4_60.jsonl | cut -c1-300
grep -q '"grid": 4,' gpurun_out/g/c4_60.jsonl && { echo "BAD GRID"; exit 1; }
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 500 2>/dev/null | tee gpurun_out/g/c4_500.jsonl | cut -c1-500
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-grad 2>/dev/null | tee gpurun_out/g/c4_grad_100.jsonl | cut -c1-500
rm -f gpurun_out/g/c4_order.jsonl
for o in rcm random; do
  timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 60 --c4-order $o 2>/dev/null | tee -a gpurun_out/g/c4_order.jsonl | cut -c1-300
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlgk_solve -s 1 -c 1 -f -o gpurun_out/g/prof_c4 python tools/profile_c4.py --n-graphs 24 > gpurun_out/g/ncu.log 2>&1
echo "ncu rc=$?"
