# usage (GPU box): bash tools/sweep_block.sh  -- A/B sweep of small-kernel build options
run() { tag=$1; shift; python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-reference-gpu "$@" > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d['value'], d['roofline']['kernel'], d['roofline'].get('mufu_model_frac'))
" gpurun_out/sw_$tag.json $tag; }
run default
run noskip --nvrtc-extra=-DGDB_SKIP_LAST_ROW=0
run default2
