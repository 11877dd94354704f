#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -k "reorder or first_need" 2>&1 | grep -v "arn" | tail -n 12
