#!/bin/bash
mkdir -p gpurun_out/b
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_reference_frontend.py tests/test_gpr.py -q -m gpu > gpurun_out/b/pytest_new.log 2>&1
echo "new tests rc=$?" > gpurun_out/b/status.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/b/bench_n1.json 2> gpurun_out/b/bench_n1.err
echo "bench rc=$?" >> gpurun_out/b/status.txt
tail -n 15 gpurun_out/b/pytest_new.log
cat gpurun_out/b/status.txt
