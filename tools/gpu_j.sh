#!/bin/bash
O=gpurun_out/j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_reference_frontend.py tests/test_gpu_pipeline.py tests/test_gpu_reference_device.py -q -m gpu > $O/pytest.log 2>&1
echo "tests rc=$?"; tail -n 2 $O/pytest.log
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 2>/dev/null | cut -c1-330
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 500 2>/dev/null | tee $O/c4_500.jsonl | cut -c1-330
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-grad 2>/dev/null | tee $O/c4_grad.jsonl | cut -c1-330
for n in 4 6 12; do
  GDB_PIPELINE_LAUNCHES=$n timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu --parity-samples 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('launches $n: value %.4g e2e %.4g e2e_ms %.2f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
timeout 120 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlgk_solve -s 1 -c 1 -f -o $O/prof_c4 python tools/profile_c4.py --n-graphs 24 > $O/ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
