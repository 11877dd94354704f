#!/bin/bash
O=gpurun_out/m; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nodal" > $O/pytest.log 2>&1
echo "nodal tests rc=$?"; tail -n 15 $O/pytest.log | cut -c1-220
python tools/bench_nodal_grad.py 300 2>/dev/null | tail -1
GDB_FORCE_GENERAL=1 python tools/bench_nodal_grad.py 300 2>/dev/null | tail -1
