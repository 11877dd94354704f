#!/bin/bash
# Build the in-tree library HERE (the .so travels with the snapshot), then run
# a command on the GPU box:  tools/gpurun.sh [gpurun options] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "from graphdot_b200.csrc import build; build.build_library()"
exec /usr/local/graft/bin/gpurun "$@"
