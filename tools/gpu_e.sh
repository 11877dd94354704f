#!/bin/bash
mkdir -p gpurun_out/e
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlgk_solve -s 1 -c 1 -f -o gpurun_out/e/prof_c4 python tools/profile_c4.py --n-graphs 16 > gpurun_out/e/ncu.log 2>&1
echo "ncu rc=$?"
tail -n 5 gpurun_out/e/ncu.log
