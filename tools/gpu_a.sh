#!/bin/bash
# round-2 GPU check A: new pipeline tests, full gpu suite, bench N=1, C5 on one GPU
mkdir -p gpurun_out/a
{ df -h /dev/shm /tmp; nproc; free -g; nvidia-smi -L; } > gpurun_out/a/box.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q > gpurun_out/a/pytest_pipeline.log 2>&1
echo "pipeline rc=$?" >> gpurun_out/a/box.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pipeline.py > gpurun_out/a/pytest_gpu.log 2>&1
echo "gpu rc=$?" >> gpurun_out/a/box.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a/bench_n1.json 2> gpurun_out/a/bench_n1.err
echo "bench rc=$?" >> gpurun_out/a/box.txt
timeout 600 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/a/bench_c5_n1.json 2> gpurun_out/a/bench_c5_n1.err
echo "c5 rc=$?" >> gpurun_out/a/box.txt
tail -5 gpurun_out/a/pytest_pipeline.log gpurun_out/a/pytest_gpu.log
cat gpurun_out/a/box.txt
