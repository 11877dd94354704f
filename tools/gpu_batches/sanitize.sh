#!/bin/bash
# compute-sanitizer memcheck + racecheck over the small-, large-pair and general kernels
O=gpurun_out/sanitize; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_target.py --large > $O/memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -n 4 $O/memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_target.py --large > $O/racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -n 4 $O/racecheck.txt
