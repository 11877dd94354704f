#!/bin/bash
# C4 large-pair (cluster) kernel: parity, sanitizer, throughput by cluster size
mkdir -p gpurun_out/d
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -k "c4" > gpurun_out/d/pytest_c4.log 2>&1
echo "c4 parity rc=$?" > gpurun_out/d/status.txt
tail -n 25 gpurun_out/d/pytest_c4.log
for C in 4 2 8; do
  GDB_CLUSTER=$C timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 60 > gpurun_out/d/c4_cluster$C.jsonl 2> gpurun_out/d/c4_cluster$C.err
  echo "cluster $C rc=$?" >> gpurun_out/d/status.txt
  cat gpurun_out/d/c4_cluster$C.jsonl
done
GDB_FORCE_GENERAL=1 timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 60 > gpurun_out/d/c4_general.jsonl 2> gpurun_out/d/c4_general.err
cat gpurun_out/d/c4_general.jsonl
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 40 --c4-grad > gpurun_out/d/c4_grad.jsonl 2> gpurun_out/d/c4_grad.err
cat gpurun_out/d/c4_grad.jsonl
cat gpurun_out/d/status.txt
