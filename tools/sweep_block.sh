# usage (GPU box): bash tools/sweep_block.sh  -- A/B sweep of bench / small-kernel options
run() { tag=$1; shift; python bench.py --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-reference-gpu "$@" > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['e2e']['value'], d['gpu_launches'])
" gpurun_out/sw_$tag.json $tag; }
run t64
run t128 --tile-rows 128
run t256 --tile-rows 256
run t32 --tile-rows 32
