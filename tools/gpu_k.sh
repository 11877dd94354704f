#!/bin/bash
O=gpurun_out/k; mkdir -p $O
run() { echo "== $*"; env "$@" timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('pairs_per_s','kernel_ms','grid','smem_bytes','hbm_frac')})"; }
run GDB_X=1
run GDB_CLUSTER=2
run GDB_CLUSTER=1
run GDB_NVRTC_EXTRA="-DGDB_LELL=8"
run GDB_NVRTC_EXTRA="-DGDB_LBLOCK=1024 -DGDB_LMINB=1"
run GDB_NVRTC_EXTRA="-DGDB_LBLOCK=256 -DGDB_LMINB=4"
run GDB_NVRTC_EXTRA="-DGDB_LBLOCK=256 -DGDB_LMINB=3" GDB_CLUSTER=2
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-grad 2>/dev/null | tee $O/c4_grad.jsonl | cut -c1-330
