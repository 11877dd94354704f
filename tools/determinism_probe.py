#!/usr/bin/env python
"""Where does the Jacobian of the large-pair kernel differ between identical calls?"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import make_config_kernel, newman_watts_strogatz  # noqa: E402

G = [newman_watts_strogatz(np.random.default_rng(s), n) for s, n in ((1, 41), (2, 56), (3, 64))]
variants = [dict(), dict(GDB_CLUSTER='1'), dict(GDB_LARGE_CLUSTERS='1'), dict(GDB_CLUSTER='4')]
for env in variants:
    env = dict(env, GDB_SMEM_CAP='40000')
    os.environ.update(env)
    res = []
    for k in range(5):
        be = B200Backend()
        K, dK = make_config_kernel('C4', backend=be)(G, eval_gradient=True)
        res.append((K.copy(), dK.copy(), be.last.get('cg_iterations'), be.last['grid'], be.last['smem_bytes']))
    for key in env:
        del os.environ[key]
    print(env, 'kernel', be.last['kernel'], 'grid', res[0][3], 'smem', res[0][4], 'iterations', [r[2] for r in res])
    for k in range(1, 5):
        d = res[k][1] != res[0][1]
        if d.any():
            idx = np.argwhere(d)
            print('  run', k, 'differs at', idx.tolist()[:8], 'values', res[0][1][d][:4], res[k][1][d][:4])
        else:
            print('  run', k, 'identical')
