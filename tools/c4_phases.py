#!/usr/bin/env python
"""Where the time of a C4 solve goes: kernel time as a function of the CG
iteration count (ftol) and of the node microkernel's cost."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphdot_b200.kernel.marginalized import MarginalizedGraphKernel
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend
from graphdot_b200.microkernel import (Constant, Convolution, SquareExponential, TensorProduct, KroneckerDelta)
from graphdot_b200.synthetic import make_config_graphs

G = make_config_graphs('C4', int(sys.argv[1]) if len(sys.argv) > 1 else 60)
be = B200Backend()
conv = TensorProduct(feat=Convolution(SquareExponential(1.0)))
edge = TensorProduct(length=SquareExponential(0.2))
for name, kn, ke in (('conv', conv, edge),):
    for ftol in (1e-8, 1e-4, 1e-2, 1.0):
        k = MarginalizedGraphKernel(kn, ke, q=0.05, ftol=ftol, backend=be)
        k(G)
        k(G)
        L = be.last
        print(json.dumps(dict(node=name, ftol=ftol, kernel_ms=L['kernel_ms'], kernel=L['kernel'],
                              it_per_pair=L['cg_iterations'] / L['n_jobs'], grid=L['grid'])), flush=True)
