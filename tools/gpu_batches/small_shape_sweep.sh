#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/sweep_small.py "$@" 2>&1 | grep -v arn | tee gpurun_out/sweep_small.jsonl
