# usage (GPU box): bash tools/sweep_block.sh  -- A-B sweep of bench / small-kernel build options
run() { tag=$1; shift; python bench.py --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-reference-gpu "$@" > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'])
" gpurun_out/sw_$tag.json $tag; }
run default
# experimental, unmeasured at the end of round 1 (run the parity tests with the same option first:
#   GDB_NVRTC_EXTRA is not a thing -- use B200Backend(nvrtc_extra=[...]) in a test or bench --nvrtc-extra)
run pipeline --nvrtc-extra=-DGDB_K1_PIPELINE=1
run default2
