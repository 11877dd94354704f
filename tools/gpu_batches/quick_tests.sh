#!/bin/bash
# quick GPU check of selected tests: bash tools/gpu_batches/quick_tests.sh "<pytest -k expression>" [files...]
K="$1"; shift
timeout 1200 python -m pytest ${@:-tests} -q -m gpu -k "$K" > gpurun_out/quick_tests.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error" gpurun_out/quick_tests.log | tail -n 8
