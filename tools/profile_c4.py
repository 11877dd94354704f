#!/usr/bin/env python
"""Short ncu target for the large-pair (cluster) kernel: symmetric Gram of a
few C4 graphs (200-500 nodes, Convolution node kernel).

    ncu --set full --clock-control none --import-source on \
        -k regex:mlgk_solve -s 1 -c 1 -o gpurun_out/prof_c4 \
        python tools/profile_c4.py [--n-graphs 16] [--grad]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n-graphs', type=int, default=16)
ap.add_argument('--launches', type=int, default=2)
ap.add_argument('--grad', action='store_true')
args = ap.parse_args()
G = make_config_graphs('C4', args.n_graphs)
be = B200Backend()
kernel = make_config_kernel('C4', backend=be)
for k in range(args.launches):
    kernel(G, eval_gradient=args.grad)
    print(k, be.last['kernel_ms'], 'ms', be.last['n_jobs'], 'pairs',
          be.last['kernel'], 'grid', be.last['grid'], 'smem',
          be.last['smem_bytes'], 'iters', be.last['cg_iterations'])
