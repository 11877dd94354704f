#!/bin/bash
# round-2 closing batch after the per-kernel NVRTC modules: tests, bench, first-call time
O=gpurun_out/final3; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/status.txt; tail -n 3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" | tee -a $O/status.txt; cat $O/bench_n1.json
timeout 300 python tools/first_call.py 2>&1 | grep -v arn | tee $O/first_call.jsonl
