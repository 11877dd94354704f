#!/bin/bash
O=gpurun_out/lb256; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -k "large or c4 or C4 or reorder or first_need or cluster" > $O/pytest_large.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" $O/pytest_large.log | tail -n 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlgk_solve -s 1 -c 1 -f -o $O/prof_c4 python tools/profile_c4.py --n-graphs 24 > $O/ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
timeout 900 python tools/bench_configs.py --only C4 --c4-graphs 500 2>&1 | grep "^{" > $O/c4_all500.jsonl; cat $O/c4_all500.jsonl
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-grad 2>&1 | grep "^{" >> $O/c4_all500.jsonl; tail -n 1 $O/c4_all500.jsonl
