#!/usr/bin/env python
"""Time breakdown of the device-resident call that bench.py's `value` times:
Normalization(kernel).device_gram(G, eval_gradient=True) on the 2000-molecule C3 set."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphdot_b200.kernel.fix import Normalization  # noqa: E402
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel  # noqa: E402

G = make_config_graphs('C2', 2000)
be = B200Backend()
norm = Normalization(make_config_kernel('C3', backend=be))
for k in range(5):
    be.reset_totals()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    K, dK = norm.device_gram(G, eval_gradient=True)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'call {k}: events {e0.elapsed_time(e1):.2f} ms, host returns after {1e3 * (t1 - t0):.2f} ms, '
          f'synced after {1e3 * (t2 - t0):.2f} ms; kernels {be.totals["kernel_ms"]:.2f} ms in '
          f'{be.totals["launches"]} launches', flush=True)
os.environ['GDB_TRACE'] = '1'
norm.device_gram(G, eval_gradient=True)
torch.cuda.synchronize()
del os.environ['GDB_TRACE']
cProfile.run('norm.device_gram(G, eval_gradient=True); torch.cuda.synchronize()', '/tmp/value.prof')
pstats.Stats('/tmp/value.prof').sort_stats('cumtime').print_stats(22)
