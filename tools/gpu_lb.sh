#!/bin/bash
# large-pair kernel: threads per CTA : resident CTAs asked of ptxas : cluster size
mkdir -p gpurun_out
for cfg in "${@:-512:2:2}"; do
  IFS=: read lb mb cl <<< "$cfg"
  GDB_CLUSTER=${cl:-2} GDB_NVRTC_EXTRA="-DGDB_LBLOCK=$lb -DGDB_LMINB=$mb" timeout 600 python tools/bench_configs.py --only C4 --c4-graphs ${C4N:-100} 2>&1 | grep -v "arn" | tail -n 1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$cfg', round(d['pairs_per_s']), d['grid'], d['smem_bytes'], d['kernel'])"
done | tee gpurun_out/c4_lblock.txt
