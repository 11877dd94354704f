#!/usr/bin/env python
"""Bit-reproducibility of every solver kernel: the same call three times in fresh back ends
(Gram + Jacobian) must return identical bits.  One line per kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import (make_config_graphs, make_config_kernel,  # noqa: E402
                                     newman_watts_strogatz)


def runs(cfg, G, n=3, **env):
    out = []
    for k in range(n):
        os.environ.update(env)
        be = B200Backend()
        K, dK = make_config_kernel(cfg, backend=be)(G, eval_gradient=True)
        for key in env:
            del os.environ[key]
        out.append((K.copy(), dK.copy(), be.last['kernel']))
    return out


cases = [('small', 'C3', make_config_graphs('C2', 60), {}),
         ('large', 'C4', make_config_graphs('C4', 6), {}),
         ('large, mid-size graphs', 'C4',
          [newman_watts_strogatz(np.random.default_rng(s), n) for s, n in ((1, 41), (2, 56), (3, 64))],
          {'GDB_SMEM_CAP': '40000'}),
         ('general', 'C4', make_config_graphs('C4', 6), {'GDB_FORCE_GENERAL': '1'}),
         ('general, mid-size graphs', 'C4',
          [newman_watts_strogatz(np.random.default_rng(s), n) for s, n in ((1, 41), (2, 56), (3, 64))],
          {'GDB_FORCE_GENERAL': '1'})]
ok = True
for name, cfg, G, env in cases:
    r = runs(cfg, G, **env)
    same_K = all(np.array_equal(r[0][0], x[0]) for x in r[1:])
    same_dK = all(np.array_equal(r[0][1], x[1]) for x in r[1:])
    worst = max(float(np.abs(x[1] - r[0][1]).max() / np.abs(r[0][1]).max()) for x in r[1:])
    print(f'{name}: kernel {r[0][2]}, Gram identical {same_K}, Jacobian identical {same_dK} '
          f'(max rel diff {worst:.3g})', flush=True)
    ok &= same_K and same_dK
print('determinism', 'ok' if ok else 'BROKEN')
