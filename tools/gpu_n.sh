#!/bin/bash
O=gpurun_out/n; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -k "pipelined or normalized_public" > $O/pytest.log 2>&1
echo "tests rc=$?"; tail -n 2 $O/pytest.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu --parity-samples 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g e2e_ms %.2f launches %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['launches_per_step']))"; }
run GDB_X=1
run GDB_PIPELINE_LAUNCHES=8
run GDB_PIPELINE_LAUNCHES=3
run GDB_PIPELINE_LAUNCHES=5
