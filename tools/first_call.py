"""First-call latency of the public API in a fresh process: NVRTC compile of the kernels a
workload needs (general + small-pair concurrently; the large-pair kernel only when a graph set
has large pairs), graph packing and upload, first launch.  One JSON line per workload."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel  # noqa: E402

for cfg, n in (('C3', 300), ('C4', 12)):
    G = make_config_graphs(cfg, n)
    be = B200Backend()
    kernel = make_config_kernel(cfg, backend=be)
    t0 = time.perf_counter()
    kernel(G, eval_gradient=True)
    t1 = time.perf_counter()
    kernel(G, eval_gradient=True)
    t2 = time.perf_counter()
    info = [be.program_info(p) for p in be._programs.values()]
    print(json.dumps(dict(config=cfg, graphs=n, kernel=be.last['kernel'],
                          first_call_s=t1 - t0, second_call_s=t2 - t1,
                          compile_ms=[i.compile_ms for i in info],
                          num_regs_large=[i.num_regs_large for i in info])),
          flush=True)
