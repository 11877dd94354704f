#!/bin/bash
mkdir -p gpurun_out/f
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -k "c4" > gpurun_out/f/pytest_c4.log 2>&1
echo "c4 parity rc=$?"; tail -n 2 gpurun_out/f/pytest_c4.log
timeout 120 python tools/bench_configs.py --only C4 --c4-graphs 60 2>/dev/null | cut -c1-330
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 200 2>/dev/null | cut -c1-330
GDB_CLUSTER=2 timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 200 2>/dev/null | cut -c1-330
