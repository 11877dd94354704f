#!/usr/bin/env python
"""Launch-shape sweep of the small-pair kernel on the C3 workload of bench.py (device-resident
normalized Gram + Jacobian of 2000 molecules): threads per pair x resident threads per SM asked
of ptxas (GDB_SMALL_THREADS, which sets the register cap).  One JSON line per shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.fix import Normalization  # noqa: E402
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel  # noqa: E402

shapes = sys.argv[1:] or ['0:640', '192:768', '256:768', '256:1024', '192:960',
                          '160:800', '128:768', '0:640']
G = make_config_graphs('C2', 2000)
ref = None
for shape in shapes:
    block, threads = (int(v) for v in shape.split(':'))
    os.environ['GDB_SMALL_THREADS'] = str(threads)
    be = B200Backend(block_size=block or None)
    norm = Normalization(make_config_kernel('C3', backend=be))
    ms = []
    for k in range(4):
        K, dK = norm.device_gram(G, eval_gradient=True)
        ms.append(be.last['kernel_ms'])
    info = [be.program_info(p) for p in be._programs.values()]
    K = K.cpu().numpy()
    if ref is None:
        ref = K
    print(json.dumps(dict(block=block, threads=threads, kernel=be.last['kernel'],
                          grid=be.last['grid'], smem=be.last['smem_bytes'],
                          regs=[i.num_regs_small for i in info],
                          kernel_ms=min(ms[1:]), all_ms=ms,
                          mpairs_per_s=be.last['n_jobs'] / min(ms[1:]) / 1e3,
                          max_rel_vs_first=float(abs(K - ref).max() / abs(ref).max()))),
          flush=True)
