#!/bin/bash
# small-kernel A/B hooks on the contract bench (value only)
mkdir -p gpurun_out/h
run() { name=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu --parity-samples 40 "$@" > gpurun_out/h/$name.json 2> gpurun_out/h/$name.err; python -c "
import json,sys
d=json.load(open('gpurun_out/h/$name.json'))
print('$name', 'value %.4g e2e %.4g kernel_ms %.2f parity %s' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['parity']['ok']))"; }
run base
run pipeline --nvrtc-extra=-DGDB_K1_PIPELINE=1
run unroll2 --nvrtc-extra=-DGDB_K1_UNROLL=2
run base2
