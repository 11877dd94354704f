#!/bin/bash
# large-pair kernel A/B over NVRTC defines: each argument is a quoted option string
mkdir -p gpurun_out
for cfg in "$@"; do
  GDB_NVRTC_EXTRA="$cfg" timeout 600 python tools/bench_configs.py --only C4 --c4-graphs ${C4N:-100} ${C4ARGS} 2>&1 | grep "^{\"config" | tail -n 1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$cfg', round(d['pairs_per_s']), d['grid'], d['smem_bytes'], d['kernel'])"
done | tee gpurun_out/c4_ab.txt
