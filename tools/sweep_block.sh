# usage (GPU box): bash tools/sweep_block.sh  -- A/B sweep of small-kernel build options
run() { tag=$1; shift; python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-reference-gpu "$@" > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['roofline']['kernel'])
" gpurun_out/sw_$tag.json $tag; }
run default
run unroll2 --nvrtc-extra=-DGDB_K1_UNROLL=2
run unroll4 --nvrtc-extra=-DGDB_K1_UNROLL=4
run m6 --nvrtc-extra=-DGDB_SMALL_MINB=6
