#!/bin/bash
# large-pair / general kernel check: their tests, then C4 throughput with and without Jacobian
O=gpurun_out/large_check; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -k "large or c4 or C4 or closed_form or 560 or reorder or rare or arena or placements" > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" $O/pytest.log | tail -n 2; grep -E "^E  " $O/pytest.log | head -12
timeout 900 python tools/bench_configs.py --only C4 --c4-graphs 500 2>&1 | grep "^{\"config" > $O/c4.jsonl
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-grad 2>&1 | grep "^{\"config" >> $O/c4.jsonl
GDB_FORCE_GENERAL=1 timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 2>&1 | grep "^{\"config" >> $O/c4.jsonl
python -c "
import json
for l in open('$O/c4.jsonl'):
    d=json.loads(l); print(d['config'], d['n_graphs'], d['kernel'], round(d['pairs_per_s']), round(d['hbm_frac'],4))"
