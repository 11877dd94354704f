#!/bin/bash
# tools/gpurun_retry.sh [gpurun options] -- '<command>': retry while the pod answers busy (exit 3 / transient)
cd "$(dirname "$0")/.."
python -c "from graphdot_b200.csrc import build; build.build_library()"
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out"
  if echo "$out" | grep -q "status=transient\|retry in a few minutes"; then sleep 150; continue; fi
  exit $rc
done
