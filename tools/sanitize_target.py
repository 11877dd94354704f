#!/usr/bin/env python
"""Small target for compute-sanitizer (memcheck / racecheck / initcheck):
molecular pairs with and without Jacobian, dense graphs that exercise helper
lanes and the overflow path, nodal output.

    compute-sanitizer --tool memcheck python tools/sanitize_target.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.fix import Normalization  # noqa: E402
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import (make_config_graphs, make_config_kernel,  # noqa: E402
                                     random_labeled_graph)

rng = np.random.default_rng(5)
for slots in (2, 4):
    be = B200Backend(slots_per_lane=slots)
    kernel = make_config_kernel('C2', backend=be)
    G = make_config_graphs('C2', 12) + [random_labeled_graph(rng, n, 0.35) for n in (9, 14, 17)]
    K, dK = kernel(G, eval_gradient=True)
    assert be.last['small_kernel']
    K2 = kernel(G)
    Kn = Normalization(kernel)(G, eval_gradient=True)
    Kd = kernel(G[:4], nodal=True)
    print(slots, float(np.abs(K - K2).max()), K.shape, dK.shape, Kd.shape)
print('sanitize target done')

# round 2: the large-pair (cluster) kernel, forced onto mid-size graphs by a
# small shared-memory cap so that the sanitizer finishes quickly; Gram +
# Jacobian, symmetric and X-by-Y, against the general kernel
if '--large' in sys.argv:
    from graphdot_b200.synthetic import newman_watts_strogatz
    os.environ['GDB_SMEM_CAP'] = '40000'
    G = [newman_watts_strogatz(np.random.default_rng(s), n) for s, n in
         ((1, 41), (2, 56), (3, 64))]
    be = B200Backend()
    kernel = make_config_kernel('C4', backend=be)
    K, dK = kernel(G, eval_gradient=True)
    print('large:', be.last['kernel'], be.last['grid'], be.last['smem_bytes'])
    assert be.last['kernel'] == 'mlgk_solve_large'
    Kxy = kernel(G[:2], G[1:])
    os.environ['GDB_FORCE_GENERAL'] = '1'
    be2 = B200Backend()
    K2, dK2 = make_config_kernel('C4', backend=be2)(G, eval_gradient=True)
    assert be2.last['kernel'] == 'mlgk_solve'
    print('large vs general:', float(np.abs(K / K2 - 1).max()),
          float(np.abs(dK - dK2).max() / np.abs(dK2).max()))
    print('sanitize large target done')
