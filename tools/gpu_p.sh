#!/bin/bash
O=gpurun_out/memo; mkdir -p $O
timeout 300 python tools/value_breakdown.py 2>&1 | grep -E "^call|cumulative|device_gram|launch|function calls" | head -14
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_gpr.py tests/test_gpu_reference_frontend.py -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" $O/pytest.log | tail -n 3
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity']['ok'])"
