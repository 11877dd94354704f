# usage (GPU box): bash tools/sweep_block.sh  -- A/B sweep of the small-kernel launch shape
run() { tag=$1; shift; python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-reference-gpu "$@" > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['roofline']['kernel'])
" gpurun_out/sw_$tag.json $tag; }
run default
run m6 --nvrtc-extra=-DGDB_SMALL_MINB=6
run m4 --nvrtc-extra=-DGDB_SMALL_MINB=4
run b160_m4 --block-size 160 --nvrtc-extra=-DGDB_SMALL_MINB=4
run b96_m6 --block-size 96 --nvrtc-extra=-DGDB_SMALL_MINB=6
run b96_m5 --block-size 96 --nvrtc-extra=-DGDB_SMALL_MINB=5
run adj4 --slots-per-lane 4
