#!/bin/bash
O=gpurun_out/lb256; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -k "large or c4 or C4 or reorder or first_need or cluster" 2>&1 | grep -v "arn" | tail -n 6
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py --large > $O/sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -n 2 $O/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py --large > $O/sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -n 2 $O/sanitizer_racecheck.txt
