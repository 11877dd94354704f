#!/usr/bin/env python
"""Nodal Jacobian throughput: small-pair kernel (round 2) vs the general kernel
(GDB_FORCE_GENERAL=1) on C2 molecules."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
G = make_config_graphs('C2', n)
be = B200Backend()
k = make_config_kernel('C3', backend=be)
for rep in range(2):
    t0 = time.perf_counter()
    R, dR = k(G, nodal=True, eval_gradient=True)
    dt = time.perf_counter() - t0
print(json.dumps(dict(kernel=be.last['kernel'], pairs=n * (n + 1) // 2, kernel_ms=be.last['kernel_ms'],
                      pairs_per_s=n * (n + 1) // 2 / (be.last['kernel_ms'] * 1e-3), call_s=dt,
                      cg_iterations_per_pair=be.last['cg_iterations'] / be.last['n_jobs'])))
