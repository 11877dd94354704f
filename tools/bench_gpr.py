#!/usr/bin/env python
"""One GPR training step (objective + gradient over all hyper-parameters) on
the C3 workload: device-resident path against the host path the reference
takes (Gram + Jacobian copied to the host, numpy/LAPACK on the CPU,
reference model/gaussian_process/gpr.py:259-296).

    python tools/bench_gpr.py [--n-graphs 2000] [--repeat 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.fix import Normalization  # noqa: E402
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.model.gaussian_process import GaussianProcessRegressor  # noqa: E402
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n-graphs', type=int, default=2000)
    ap.add_argument('--repeat', type=int, default=3)
    args = ap.parse_args()
    import torch
    G = make_config_graphs('C2', args.n_graphs)
    rng = np.random.default_rng(0)
    y = np.array([len(g.nodes) + 0.1 * rng.standard_normal() for g in G])
    be = B200Backend()
    kernel = Normalization(make_config_kernel('C2', backend=be))
    out = {'n_graphs': args.n_graphs, 'n_hyperparameters': len(kernel.theta)}
    vals = {}
    for name, device in (('device', 'auto'), ('host', 'cpu')):
        gpr = GaussianProcessRegressor(kernel, alpha=1e-3, normalize_y=True,
                                       device=device)
        gpr.X, gpr.y = G, y
        gpr.log_marginal_likelihood(eval_gradient=True)      # warm-up / JIT
        best = 1e99
        for _ in range(args.repeat):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            v, g = gpr.log_marginal_likelihood(eval_gradient=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[f'{name}_step_s'] = best
        out[f'{name}_kernel_ms'] = be.last['kernel_ms']
        vals[name] = (v, g)
    out['speedup'] = out['host_step_s'] / out['device_step_s']
    out['objective_rel_diff'] = abs(vals['device'][0] - vals['host'][0]) / abs(vals['host'][0])
    out['gradient_max_rel_diff'] = float(np.max(
        np.abs(vals['device'][1] - vals['host'][1]) / np.abs(vals['host'][1]).max()))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
