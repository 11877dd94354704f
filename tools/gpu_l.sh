#!/bin/bash
O=gpurun_out/l; mkdir -p $O
python tools/c4_phases.py 100 2>&1 | grep "^{"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlgk_solve -s 1 -c 1 -f -o $O/prof_c4 python tools/profile_c4.py --n-graphs 24 > $O/ncu_c4.log 2>&1
echo "ncu rc=$?"
