#!/bin/bash
# round-2 closing batch: full GPU suite, contract bench (both arms), smoke
O=gpurun_out/final4; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/status.txt; grep -E "passed|failed|error" $O/pytest_gpu.log | tail -n 3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/status.txt; tail -n 2 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" | tee -a $O/status.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_impl_reference.json 2> $O/bench_impl_reference.err
echo "reference arm rc=$?" | tee -a $O/status.txt
python -c "
import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity']['ok'], d['roofline']['frac'], d['clocks'])
r=json.load(open('$O/bench_impl_reference.json')); print(r['value'], r['cpu_baseline'])"
