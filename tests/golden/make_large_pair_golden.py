#!/usr/bin/env python
"""Known answer for a graph pair beyond BASELINE's C4 range (200-500 nodes):
560 x 530 nodes, N = 296 800, C4 kernels.  The float64 oracle
(oracle/mlgk_oracle.py, Jacobi-CG to 1e-14 at this size) needs 80 s for the
pair, so its answer is committed instead of recomputed by the GPU test.

    python tests/golden/make_large_pair_golden.py   # writes large_pair_oracle.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from graphdot_b200.synthetic import (make_config_kernel,  # noqa: E402
                                     newman_watts_strogatz)
from oracle import mlgk_oracle as oracle  # noqa: E402

SEEDS, SIZES = (21, 22), (560, 530)


def graphs():
    return [newman_watts_strogatz(np.random.default_rng(s), n)
            for s, n in zip(SEEDS, SIZES)]


if __name__ == '__main__':
    g1, g2 = graphs()
    k = make_config_kernel('C4')
    _, ko, go = oracle.solve_pair(g1, g2, k.node_kernel, k.edge_kernel, k.q,
                                  k.p, eval_gradient=True)
    out = dict(provenance='oracle/mlgk_oracle.py solve_pair, float64, '
                          'tests/golden/make_large_pair_golden.py',
               seeds=SEEDS, sizes=SIZES, config='C4', gram=float(ko),
               gradient=[float(v) for v in go])
    with open(os.path.join(HERE, 'large_pair_oracle.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(out)
