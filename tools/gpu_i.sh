#!/bin/bash
mkdir -p gpurun_out/i
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/i/pytest_gpu.log 2>&1
echo "gpu tests rc=$?"; tail -n 4 gpurun_out/i/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --parity-samples 200 > gpurun_out/i/bench_n1.json 2> gpurun_out/i/bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/i/bench_n1.json'))
print('value %.4g e2e %.4g kernel_ms %.2f frac %.4f it/pair %.2f parity %s' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['cg_iterations_per_pair'], d['parity']))
print(d['e2e'])"
