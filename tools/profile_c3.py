#!/usr/bin/env python
"""ncu target with the launch shapes of bench.py at N = 1: the device-resident
normalized Gram + Jacobian of the 2000-molecule C3 workload (one diagonal
launch + ONE launch of all 2 001 000 pairs per call).

    ncu --set full --clock-control none --import-source on \
        -k regex:mlgk_solve_small -s 3 -c 1 -o gpurun_out/prof_c3 \
        python tools/profile_c3.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.fix import Normalization  # noqa: E402
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
G = make_config_graphs('C2', n)
be = B200Backend()
norm = Normalization(make_config_kernel('C3', backend=be))
for k in range(3):
    K, dK = norm.device_gram(G, eval_gradient=True)
    print(k, be.last['kernel_ms'], 'ms', be.last['n_jobs'], 'pairs',
          be.last['kernel'], 'grid', be.last['grid'])
