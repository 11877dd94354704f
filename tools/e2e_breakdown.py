#!/usr/bin/env python
"""Host-side time breakdown of the public call that bench.py's e2e times:
Normalization(kernel)(G, eval_gradient=True) on the 2000-molecule C3 set."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphdot_b200.kernel.fix import Normalization
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel
G = make_config_graphs('C2', 2000)
be = B200Backend()
be.resend_graphs = True
norm = Normalization(make_config_kernel('C3', backend=be))
for k in range(4):
    t0 = time.perf_counter()
    K, dK = norm(G, eval_gradient=True, timing=(k == 3))
    dt = time.perf_counter() - t0
    print(f'call {k}: {dt * 1e3:.2f} ms  kernel {be.last["kernel_ms"]:.2f} ms  launches {be.last["n_launches"]}')
import cProfile, pstats
cProfile.run('norm(G, eval_gradient=True)', '/tmp/e2e.prof')
pstats.Stats('/tmp/e2e.prof').sort_stats('cumtime').print_stats(18)
