#!/usr/bin/env python
"""Short ncu target: a few row-block tile launches of the C3 workload
(normalized-Gram tiles with Jacobian, 2000 synthetic molecules).

    ncu --set full --clock-control none -k regex:mlgk_solve -s 4 -c 2 \
        -o gpurun_out/prof python tools/profile_target.py [--no-grad] [--config C4]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.kernel.marginalized._tiles import GramTileWorker  # noqa: E402
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C2')
ap.add_argument('--n-graphs', type=int, default=2000)
ap.add_argument('--rows', type=int, default=64)
ap.add_argument('--launches', type=int, default=6)
ap.add_argument('--block-size', type=int, default=0)
ap.add_argument('--slots-per-lane', type=int, default=0)
ap.add_argument('--no-grad', action='store_true')
ap.add_argument('--nvrtc-extra', default='')
args = ap.parse_args()

G = make_config_graphs(args.config, args.n_graphs)
be = B200Backend(block_size=args.block_size or None,
                 slots_per_lane=args.slots_per_lane or None,
                 nvrtc_extra=args.nvrtc_extra.split())
kernel = make_config_kernel(args.config, backend=be)
w = GramTileWorker(kernel, G, be, eval_gradient=not args.no_grad,
                   max_rows=args.rows)
for k in range(args.launches):
    w.run_tile(0, args.rows, keep_on_device=True)
    print(k, be.last['kernel_ms'], 'ms', be.last['n_jobs'], 'pairs',
          'small' if be.last['small_kernel'] else 'general',
          'grid', be.last['grid'], 'smem', be.last['smem_bytes'])
