#!/bin/bash
mkdir -p gpurun_out
run() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"config'):
        d=json.loads(l); print('$1', round(d['pairs_per_s']), d['grid'], d['smem_bytes'], d['kernel'], d.get('order'))"; }
for lb in 512 256; do
  export GDB_NVRTC_EXTRA="-DGDB_LBLOCK=$lb"
  timeout 600 python tools/bench_configs.py --only C4 --c4-graphs 500 2>&1 | run "$lb all500"
  timeout 600 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-grad 2>&1 | run "$lb grad100"
  timeout 600 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-order random 2>&1 | run "$lb random100"
done | tee gpurun_out/c4_lblock2.txt
