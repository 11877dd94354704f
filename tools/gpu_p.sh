#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -k "mixed or rare" 2>&1 | grep -v "arn" | tail -n 12
