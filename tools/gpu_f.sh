#!/bin/bash
mkdir -p gpurun_out/f
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -q -k "c4 or vario" > gpurun_out/f/pytest_c4.log 2>&1
echo "c4+vario parity rc=$?"
tail -n 3 gpurun_out/f/pytest_c4.log
python tools/c4_phases.py 60 2>&1 | grep "^{"
