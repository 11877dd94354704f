#!/bin/bash
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --parity-samples 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g e2e_ms %.2f launches %s parity %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['launches_per_step'], d['parity']['ok']))"
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpr.py -q -m gpu 2>&1 | tail -n 2
