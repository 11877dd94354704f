#!/usr/bin/env python
"""Benchmark of the MLGK hot path (BASELINE.json metric: graph-pair MLGK
solves/sec, Gram + gradient).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm

Workloads (``config.workload``):

* N = 1 -- BASELINE config C3 (= C2 + ``eval_gradient=True``): the normalized
  Gram matrix AND its Jacobian over all 5 hyper-parameters of the 2000
  synthetic molecules (2 001 000 graph pairs per step).
* N > 1 -- BASELINE config C5 at its stated, FIXED size (strong scaling): the
  10 000 x 10 000 off-diagonal block X x Y of the 20 000-molecule seed-5005
  set, normalized Gram + Jacobian = 10^8 pairs per step.  The job rectangle
  is cut into column tiles that the ranks pull from one dynamic queue (a
  counter in the torch.distributed store); rank 0 packs the graphs once and
  broadcasts the packed set; every rank copies its finished tiles straight
  into ONE page-locked shared-memory host matrix, so rank 0 ends up holding
  the assembled result with no collective on the data path.
  (``--workload c5`` runs the same job on one GPU.)

A "step" is one pass over all pairs.

``value``   pairs/s with graphs resident in HBM and the complete result left in
            device memory (CUDA events on the launching stream, max over ranks).
``e2e``     pairs/s through the public call with HOST buffers.  N = 1:
            ``Normalization(kernel)(G, eval_gradient=True)`` -- every step
            re-sends the packed graphs (pinned) host->device and returns the
            float64 Gram + Jacobian as numpy arrays.  N > 1: the tile worker;
            every step re-sends the graphs and lands all tiles in the shared
            host matrix.
``parity``  after the timed loops, entries of the e2e result sampled at random
            are compared with the float64 CPU oracle (oracle/mlgk_oracle.py).
``roofline`` achieved FP32 FLOP/s of the solver kernel (algorithmic flops of
            SURVEY.md 8(d) with the iteration counts the engine reports) against
            the FP32 pipe peak derived from the measured SM clock.
``cpu_baseline`` / ``--impl reference``: the float64 dense Kronecker solve
            (oracle/mlgk_oracle.py, a port of the reference's own CPU oracle,
            reference test/kernel/marginalized/test_kernel.py:20-68) on all
            host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'graph-pair MLGK solves/sec (Gram + gradient)'
UNIT = 'pairs/s'
# FP32 ops of the microkernel value / Jacobian expressions of the C2/C3
# kernels (a transcendental counts 1), SURVEY.md 8(d)
F_E, F_DE = 6, 3           # edge: w1*w2, sub, sq, scale, exp, mul | d/dls
F_V, F_DV = 7, 6           # node: cmp, select, sub, sq, scale, exp, mul
C5_NX = C5_NY = 10000
N_CPU_SAMPLE = 96          # graphs whose pairs the CPU arm times


def pick_workload(args, world):
    return (args.workload or ('c3' if world == 1 else 'c5')).lower()


def workload_config(workload):
    base = {
        'node_kernel': 'TensorProduct(element=KroneckerDelta(0.5), '
                       'x=SquareExponential(1.0))',
        'edge_kernel': 'TensorProduct(length=SquareExponential(0.1))',
        'q': 0.05, 'outputs': 'normalized Gram + Jacobian over all 5 '
                              'hyper-parameters (p, q, h, l_node, l_edge)',
        'l2': 'flushed between timed steps (256 MiB write)',
    }
    if workload == 'c3':
        n = 2000
        return dict(base, workload='C3 = C2 + eval_gradient: 2000 synthetic '
                    'molecular graphs (16-24 nodes), symmetric normalized '
                    'Gram + Jacobian', n_graphs=n,
                    pairs_per_step=n * (n + 1) // 2)
    return dict(base, workload='C5: 10000 x 10000 off-diagonal block X x Y of '
                'the 20000-molecule seed-5005 set (16-24 nodes), fixed size '
                '(strong scaling), normalized Gram + Jacobian',
                n_graphs=C5_NX + C5_NY, pairs_per_step=C5_NX * C5_NY,
                sharding='column tiles of the job rectangle from a dynamic '
                         'queue, graphs replicated (packed once, broadcast)')


# ---------------------------------------------------------------------------
# CPU arm (oracle port), also the cpu_baseline of the GPU arm and the parity
# checker
# ---------------------------------------------------------------------------
class _NoBackend:
    """Placeholder so the CPU arm can build the kernel object (for its
    hyper-parameters) without touching the CUDA library."""
    def __new__(cls):
        from graphdot_b200.kernel.marginalized._backend import Backend

        class Dummy(Backend):
            def __call__(self, *a, **k):
                raise RuntimeError('CPU arm has no engine')
        return Dummy()


def _cpu_kernel():
    k = _cpu_kernel.kernel
    if k is None:
        from graphdot_b200.synthetic import make_config_kernel
        k = _cpu_kernel.kernel = make_config_kernel('C3',
                                                    backend=_NoBackend())
    return k


_cpu_kernel.kernel = None


def _cpu_pair(args):
    """(K, dK/dtheta) of one graph pair by the float64 oracle."""
    from oracle import mlgk_oracle as oracle
    g1, g2 = args
    k = _cpu_kernel()
    out = oracle.solve_pair(g1, g2, k.node_kernel, k.edge_kernel, k.q, k.p,
                            eval_gradient=True)
    return out[1], out[2]


def _pool():
    import multiprocessing as mp
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    return mp.get_context('fork').Pool(os.cpu_count() or 1)


def cpu_sample_pairs(workload):
    """The bounded sample the CPU arm times: pairs among the first 96 graphs
    of the workload's set (upper triangle for C3, X[:96] x Y[:48] for C5)."""
    from graphdot_b200.synthetic import make_config_graphs
    if workload == 'c3':
        G = make_config_graphs('C2', N_CPU_SAMPLE)
        return [(G[i], G[j]) for i in range(len(G))
                for j in range(i, len(G))], \
            (f'upper triangle of the first {N_CPU_SAMPLE} graphs of the C2 '
             'set')
    G = make_config_graphs('C5', N_CPU_SAMPLE + N_CPU_SAMPLE // 2)
    X, Y = G[:N_CPU_SAMPLE], G[N_CPU_SAMPLE:]
    return [(a, b) for a in X for b in Y], \
        (f'{N_CPU_SAMPLE} x {N_CPU_SAMPLE // 2} block of the first '
         f'{len(G)} graphs of the C5 set')


def cpu_sample_rate(workload, pool, repeats=1):
    pairs, what = cpu_sample_pairs(workload)
    cores = os.cpu_count() or 1
    pool.map(_cpu_pair, pairs[:cores])           # warm the workers
    t0 = time.perf_counter()
    for _ in range(repeats):
        pool.map(_cpu_pair, pairs, chunksize=max(1, len(pairs) // (8 * cores)))
    dt = (time.perf_counter() - t0) / repeats
    return len(pairs) / dt, cores, len(pairs), dt, what


def run_reference(args, rank, world):
    if rank != 0:
        return
    workload = pick_workload(args, world)
    pool = _pool()
    times = []
    try:
        for s in range(args.warmup + args.steps):
            rate, cores, n_pairs, dt, what = cpu_sample_rate(workload, pool)
            if s >= args.warmup:
                times.append(dt)
    finally:
        pool.close()
        pool.join()
    dt = float(np.mean(times))
    value = n_pairs / dt
    sample = (f'{n_pairs} pairs per step ({what}), float64 dense Kronecker '
              'solve + adjoint Jacobian (oracle/mlgk_oracle.py); the rate is '
              'measured on this sample, not on the full pairs_per_step')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak' if workload == 'c3' else 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(workload),
        'timed_pairs_per_step': n_pairs,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def check_parity(pool, G, pairs, K, dK, diag_cache=None):
    """Compare sampled normalized entries K[r, c], dK[r, c, :] -- the pair
    (graph a, graph b) listed as (r, c, a, b) -- with the float64 oracle:
    raw pair solves + the quotient rule of reference kernel/fix.py:46-73."""
    need = sorted({a for _, _, a, b in pairs} | {b for _, _, a, b in pairs})
    selfs = dict(zip(need, pool.map(_cpu_pair, [(G[a], G[a]) for a in need],
                                    chunksize=8)))
    cross = pool.map(_cpu_pair, [(G[a], G[b]) for _, _, a, b in pairs],
                     chunksize=8)
    err_k = err_g = 0.0
    scale_g = np.zeros(dK.shape[2])
    diff_g = np.zeros(dK.shape[2])
    for (r, c, a, b), (kab, dab) in zip(pairs, cross):
        kaa, daa = selfs[a]
        kbb, dbb = selfs[b]
        kn = kab / np.sqrt(kaa * kbb)
        dn = dab / np.sqrt(kaa * kbb) - 0.5 * kn * (daa / kaa + dbb / kbb)
        err_k = max(err_k, abs(K[r, c] - kn) / abs(kn))
        diff_g = np.maximum(diff_g, np.abs(dK[r, c, :] - dn))
        # scale: the raw-gradient term of the quotient rule (the plane of the
        # starting probability cancels to exactly 0 after normalization)
        scale_g = np.maximum(scale_g, np.abs(dab) / np.sqrt(kaa * kbb))
    err_g = float(np.max(diff_g / scale_g))
    return {'max_rel_gram': float(err_k), 'max_rel_grad': err_g,
            'n': len(pairs), 'tolerance': {'gram': 1e-5, 'grad': 1e-4},
            'ok': bool(err_k < 1e-5 and err_g < 1e-4),
            'against': 'oracle/mlgk_oracle.py (float64 dense solve), '
                       'gradient error per plane relative to the largest '
                       'sampled |dK_ij| / sqrt(K_ii K_jj) of that plane'}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,'
         'clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '100',
                 '-i', str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['n/a']}
        self.proc.terminate()
        out, _ = self.proc.communicate(timeout=10)
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for line in out.splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        busy = [c for c in sm if smax and c > 0.5 * smax] or sm
        return {'sm_mhz': float(np.median(busy)) if busy else None,
                'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


def reference_gpu_rate(n, q):
    """pairs/s of the REFERENCE's own device code (unmodified template.cu +
    graphdot/cpp compiled for sm_100a, oracle/_ref/c3_grad) on this GPU for the
    same graphs: symmetric Gram + Jacobian of the first n graphs, one warm-up
    and one timed launch, device time of graph_kernel_solver."""
    from oracle import ref_device
    if not ref_device.available('c3_grad'):
        return {'unavailable': 'oracle/_ref/c3_grad not built'}
    try:
        ref = ref_device.RefDeviceSolver('c3_grad')
        n = min(n, ref.n_graphs)
        jobs = ref_device.triu_jobs(n)
        ref.solve(ref_device.triu_jobs(64), q, n=n)
        K, dK, ms = ref.solve(jobs, q, n=n)
        return {'value': len(jobs) / (ms * 1e-3), 'unit': UNIT,
                'kernel_ms': ms, 'pairs': len(jobs), 'n_graphs': n,
                'what': 'reference graph_kernel_solver (grid SMs x 8, block '
                        '128, --maxrregcount=64) launched through the CUDA '
                        'driver API; raw (un-normalized) Gram + Jacobian'}
    except Exception as e:       # the comparator must never break the bench
        return {'unavailable': f'{type(e).__name__}: {e}'}


def algorithmic_flops(products, vec_elems, sum_N, sum_nnzx, nJ):
    """FP32 flops of the solver per SURVEY.md 8(d): matvec
    W_mv = nnzx (2 + F_e) + 2N per application, plus 13 N per CG iteration,
    setup N (F_v + 4) per solve, and the Jacobian sweep."""
    matvec = products * (2 + F_E) + 2 * vec_elems
    cg = matvec + 13 * vec_elems
    setup = sum_N * (F_V + 4) * 2
    jac = sum_nnzx * (2 + F_E + F_DE) + sum_N * (F_DV + 8 + 2 * nJ)
    return {'matvec': matvec, 'total': cg + setup + jac}


class SharedHostMatrix:
    """ONE host copy of the (nx, ny) Gram and (nx, ny, nJ) Jacobian, Fortran
    ordered float32, in a shared-memory file that every rank maps and
    page-locks: ranks copy their tiles into it directly ("gathering Gram tiles
    back to host" without a collective)."""

    def __init__(self, tag, nx, ny, nJ, rank, barrier):
        from graphdot_b200 import native
        self.nbytes = nx * ny * (1 + nJ) * 4
        d = '/dev/shm'
        try:
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize < self.nbytes * 1.05:
                d = '/tmp'
        except OSError:
            d = '/tmp'
        self.path = os.path.join(d, f'gdb_result_{tag}.bin')
        self.rank = rank
        if rank == 0:
            with open(self.path, 'wb') as f:
                f.truncate(self.nbytes)
        barrier()
        self.mm = np.memmap(self.path, dtype=np.float32, mode='r+',
                            shape=(nx * ny * (1 + nJ),))
        self.addr = self.mm.ctypes.data
        self.pinned = True
        try:
            native.check(native.load().gdb_host_register(self.addr,
                                                         self.nbytes))
        except native.NativeError:
            self.pinned = False
        self.K = self.mm[:nx * ny].reshape((nx, ny), order='F')
        self.dK = self.mm[nx * ny:].reshape((nx, ny, nJ), order='F')
        self.k_addr = self.addr
        self.dk_addr = self.addr + nx * ny * 4
        barrier()

    def close(self, barrier):
        from graphdot_b200 import native
        if self.pinned:
            native.load().gdb_host_unregister(self.addr)
        self.K = self.dK = None
        del self.mm
        barrier()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from graphdot_b200.kernel.fix import Normalization
    from graphdot_b200.kernel.marginalized._backend_b200 import (B200Backend,
                                                                 PackedGraph)
    from graphdot_b200.kernel.marginalized._tiles import (GramTileWorker,
                                                          LocalTileQueue,
                                                          StoreTileQueue,
                                                          col_tiles)
    from graphdot_b200.synthetic import make_config_graphs, make_config_kernel

    # keep stdout clean for the single JSON line (NCCL prints its version there)
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    store = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda',
                                                               local_rank))
        store = dist.distributed_c10d._get_default_store()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    workload = pick_workload(args, world)
    backend = B200Backend(device=local_rank,
                          block_size=args.block_size or None,
                          slots_per_lane=args.slots_per_lane or None,
                          nvrtc_extra=args.nvrtc_extra.split())
    kernel = make_config_kernel('C3', backend=backend)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    nJ = kernel.n_dims
    launch = {}
    first_call = {}

    if workload == 'c3':
        # ---- one GPU, symmetric 2000-graph set through the public API ---------
        n = 2000
        G = make_config_graphs('C2', n)
        total_pairs = n * (n + 1) // 2
        t0 = time.perf_counter()
        backend.pack_graphs(G)
        first_call['pack_ms'] = (time.perf_counter() - t0) * 1e3
        norm = Normalization(kernel)
        sizes = np.array([len(g.nodes) for g in G], float)
        nnz = np.array([2 * len(g.edges) for g in G], float)
        # sums over the unique pairs (i <= j) of N = n_i n_j and nnz_i nnz_j
        sum_N = (sizes.sum() ** 2 + (sizes ** 2).sum()) / 2
        sum_nnzx = (nnz.sum() ** 2 + (nnz ** 2).sum()) / 2
        result = {}

        def step_device():
            # complete normalized Gram + Jacobian left in device memory
            result['dev'] = norm.device_gram(G, eval_gradient=True)

        def step_e2e():
            backend.resend_graphs = True
            try:
                result['host'] = norm(G, eval_gradient=True)
            finally:
                backend.resend_graphs = False

        launch.update(value_path='Normalization(kernel).device_gram(G, '
                      'eval_gradient=True): 1 diagonal launch + 1 launch of '
                      'all 2 001 000 pairs, result in torch CUDA tensors',
                      e2e_path='Normalization(kernel)(G, eval_gradient=True)'
                      ': graphs H2D, diagonal launch, 4 pipelined row-block '
                      'launches, column blocks D2H + float64 collection '
                      'overlapped with the next launch')

        def sample_pairs(rng, k):
            r = rng.integers(0, n, k)
            c = rng.integers(0, n, k)
            return [(int(a), int(b), int(a), int(b)) for a, b in zip(r, c)]

        def host_result():
            return result['host']

        def cleanup():
            pass
    else:
        # ---- C5: fixed 10k x 10k block, column tiles from a dynamic queue ------
        nx, ny = C5_NX, C5_NY
        total_pairs = nx * ny
        G = None
        payload = [None]
        if rank == 0:
            G = make_config_graphs('C5', nx + ny)
            t0 = time.perf_counter()
            packed = backend.pack_graphs(G)
            first_call['pack_ms'] = (time.perf_counter() - t0) * 1e3
            base = packed[0].blob.base
            cuts = np.cumsum([0] + [p.blob.nbytes for p in packed])
            same_buf = base is not None and all(p.blob.base is base
                                                for p in packed)
            buf = (np.asarray(base) if same_buf
                   else np.concatenate([p.blob for p in packed]))
            payload = [(len(buf), cuts, [p.n_node for p in packed],
                        packed[0].key, B200Backend._layouts(G[0]))]
        if world > 1:
            # the packed set travels once over NVLink: metadata as a small
            # object, the blobs as one byte tensor (NCCL broadcast)
            t0 = time.perf_counter()
            dist.broadcast_object_list(payload, src=0)
            nbytes = payload[0][0]
            blob_t = (torch.from_numpy(buf).cuda() if rank == 0 else
                      torch.empty(nbytes, dtype=torch.uint8, device='cuda'))
            dist.broadcast(blob_t, src=0)
            if rank != 0:
                buf = blob_t.cpu().numpy()
            del blob_t
            first_call['broadcast_ms'] = (time.perf_counter() - t0) * 1e3
        _, cuts, n_nodes, key, layouts = payload[0]
        packed = [PackedGraph(buf[a:b], nn, key)
                  for a, b, nn in zip(cuts[:-1], cuts[1:], n_nodes)]
        t0 = time.perf_counter()
        gs = backend.graphset_from_packed(packed, layouts)
        first_call['upload_ms'] = (time.perf_counter() - t0) * 1e3
        worker = GramTileWorker(kernel, None, backend, eval_gradient=True,
                                max_rows=0, stream=stream.cuda_stream, nx=nx,
                                packed=gs)
        tile_cols = args.tile_cols or max(16, ny // (8 * world))
        tiles = col_tiles(ny, tile_cols)
        # the complete result of this rank's tiles stays in device memory
        Kd = torch.zeros((ny, nx), dtype=torch.float32, device='cuda')
        dKd = torch.zeros((nJ, ny, nx), dtype=torch.float32, device='cuda')
        dev = (Kd.data_ptr(), dKd.data_ptr())
        shared = SharedHostMatrix(f'{os.environ.get("MASTER_PORT", "0")}_'
                                  f'{os.getppid() if world > 1 else os.getpid()}',
                                  nx, ny, nJ, rank, barrier)
        host = (shared.k_addr, shared.dk_addr)
        sizes_x = np.array(n_nodes[:nx], float)
        sizes_y = np.array(n_nodes[nx:], float)
        hdr = [p.blob[:16].view(np.int32) for p in packed]
        nnz_all = np.array([h[2] for h in hdr], float)
        sum_N = sizes_x.sum() * sizes_y.sum()
        sum_nnzx = nnz_all[:nx].sum() * nnz_all[nx:].sum()
        step_no = [0]

        def tile_queue():
            """Dynamic tile queue: a local counter (N=1) or an atomic counter
            in the torch.distributed store shared by all ranks (N>1)."""
            key = f'tiles{step_no[0]}'
            step_no[0] += 1
            if store is None:
                return iter(LocalTileQueue(tiles))
            return iter(StoreTileQueue(store, tiles, key))

        def step_device():
            for j0, j1 in tile_queue():
                worker.run_cols(j0, j1, normalize=True, dev=dev)

        def step_e2e():
            # H2D of the packed graphs + self-similarities, then tiles whose
            # copy-back into the shared host matrix overlaps the next tile
            worker.diag(upload=True, store=True, fetch=False)
            for j0, j1 in tile_queue():
                worker.run_cols(j0, j1, normalize=True, dev=dev, host=host,
                                async_=True)
            backend.synchronize()

        worker.diag(store=True, fetch=False)
        launch.update(tile_cols=tile_cols, tiles_per_step=len(tiles),
                      host_matrix=('shared memory, page-locked by every rank'
                                   if shared.pinned else
                                   'shared memory, NOT page-locked '
                                   '(registration failed)'),
                      host_matrix_path=shared.path,
                      value_path='tile worker, result tiles left in a '
                      'full-size device matrix per rank',
                      e2e_path='tile worker: graphs H2D + diagonal launch, '
                      'every tile copied asynchronously into the shared '
                      'host matrix (copy overlaps the next tile)')

        def sample_pairs(rng, k):
            r = rng.integers(0, nx, k)
            c = rng.integers(0, ny, k)
            return [(int(a), int(b), int(a), int(nx + b))
                    for a, b in zip(r, c)]

        def host_result():
            return shared.K, shared.dK

        def cleanup():
            shared.close(barrier)

    # ---- device-resident throughput -----------------------------------------
    for _ in range(args.warmup):
        barrier()
        step_device()
    sampler = ClockSampler(local_rank)
    backend.reset_totals()
    ev = [(torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    for s in range(args.steps):
        flush.zero_()
        barrier()
        ev[s][0].record(stream)
        step_device()
        ev[s][1].record(stream)
    barrier()
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    stats = dict(backend.totals)

    # ---- end to end with host buffers ----------------------------------------
    # (at least 3 untimed steps: the pooled page-locked result buffers of the
    # public call reach their steady state -- two sets in flight -- after two)
    for _ in range(max(3, args.warmup) if workload == 'c3' else max(1, args.warmup // 2)):
        barrier()
        step_e2e()
    backend.reset_totals()
    e2e_steps = args.e2e_steps or args.steps
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        step_e2e()
        barrier()
    e2e_s = time.perf_counter() - t0
    e2e_stats = dict(backend.totals)

    # ---- reduce over ranks ------------------------------------------------------
    vec = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device='cuda')
    sums = torch.tensor([stats['kernel_ms'], stats['cg_iterations'],
                         stats['matvec_products'], stats['vector_elements'],
                         stats['launches'], stats['pairs'],
                         e2e_stats['h2d_bytes'], e2e_stats['d2h_bytes'],
                         e2e_stats['launches']],
                        dtype=torch.float64, device='cuda')
    per_rank = torch.zeros(world, dtype=torch.float64, device='cuda')
    per_rank[rank] = stats['pairs']
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s = vec.tolist()
    (kernel_ms, cg_it, products, vec_elems, launches, pairs_done, h2d,
     d2h, e2e_launches) = sums.tolist()
    n_diag = 2000 if workload == 'c3' else 0    # device_gram's diagonal solve
    assert int(pairs_done) == (total_pairs + n_diag) * args.steps, (
        pairs_done, total_pairs)
    if rank != 0:
        cleanup()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity of the assembled end-to-end result vs the CPU oracle ------------
    pool = _pool()
    try:
        Kh, dKh = host_result()
        if G is None:
            G = make_config_graphs('C5', C5_NX + C5_NY)
        pairs = sample_pairs(np.random.default_rng(12345), args.parity_samples)
        parity = check_parity(pool, G, pairs, Kh, dKh)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_sample_rate(workload, pool)
    finally:
        pool.close()
        pool.join()

    value = total_pairs * args.steps / (dev_ms * 1e-3)
    e2e_value = total_pairs * e2e_steps / e2e_s
    fl = algorithmic_flops(products / args.steps, vec_elems / args.steps,
                           sum_N, sum_nnzx, nJ)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    info = backend.device_info()
    sm_mhz = peaks.get('sm_max_mhz') or info.clock_khz / 1e3
    peak_tflops = info.sm_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    kernel_s = kernel_ms * 1e-3 / args.steps / world   # avg per rank & step
    achieved = fl['total'] / world / kernel_s / 1e12
    traffic = None      # DRAM bytes of one launch from the committed ncu capture
    traffic_note = None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
        traffic = tr[workload]['dram_bytes_per_launch']
        traffic_note = tr[workload]['note']
    except (OSError, KeyError, ValueError):
        pass
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak' if workload == 'c3' else 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(workload),
        'launch': launch,
        'e2e': {'value': e2e_value, 'unit': UNIT,
                'h2d_bytes_per_step': int(h2d / e2e_steps),
                'd2h_bytes_per_step': int(d2h / e2e_steps),
                'steps': e2e_steps, 'ms_per_step': e2e_s / e2e_steps * 1e3,
                'launches_per_step': e2e_launches / e2e_steps,
                'first_call_ms': first_call},
        'parity': parity,
        'gpu_launches': int(launches),
        'pairs_per_rank': [int(x) for x in (per_rank / args.steps).tolist()],
        'roofline': {
            'bound': 'fp32', 'achieved': achieved, 'peak': peak_tflops,
            'unit': 'TFLOP/s', 'frac': achieved / peak_tflops,
            'traffic': traffic, 'traffic_note': traffic_note,
            'peak_source': f'{info.sm_count} SMs x 128 lanes x 2 x '
                           f'{sm_mhz:.0f} MHz (measured sm_max_mhz)',
            'matvec_tflops': fl['matvec'] / world / kernel_s / 1e12,
            # SURVEY 8(d) also asks for the MUFU ceiling of an edge kernel
            # evaluated on the fly (one ex2 per product, 16 / clk / SM); this
            # kernel caches the products per pair, so it may exceed it
            'mufu_model_frac': (products / args.steps / world / kernel_s)
            / (info.sm_count * 16 * sm_mhz * 1e6),
            'kernel_ms_per_step': kernel_ms / args.steps / world,
            'cg_iterations_per_pair': cg_it / args.steps / total_pairs,
            'kernel': ('mlgk_solve_small' if backend.last.get('small_kernel')
                       else 'mlgk_solve') + f' grid={backend.last.get("grid")}'
                      f' smem={backend.last.get("smem_bytes")}',
        },
        'clocks': clocks,
    }
    if cpu is not None:
        rate, cores, n_pairs, dt, what = cpu
        line['cpu_baseline'] = {
            'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{n_pairs} pairs ({what}), {dt:.1f} s, float64 dense '
                      'Kronecker solve + adjoint Jacobian '
                      '(oracle/mlgk_oracle.py)'}
    if world == 1 and workload == 'c3' and not args.no_reference_gpu:
        line['reference_gpu'] = reference_gpu_rate(2000, kernel.q)
    cleanup()
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='', choices=['', 'c3', 'c5'],
                    help='default: c3 on one GPU, c5 on several')
    ap.add_argument('--tile-cols', type=int, default=0,
                    help='C5: columns per tile; 0 = about 8 tiles per rank')
    ap.add_argument('--block-size', type=int, default=0)
    ap.add_argument('--slots-per-lane', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=0,
                    help='0 = as many as --steps')
    ap.add_argument('--parity-samples', type=int, default=1000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-gpu', action='store_true')
    ap.add_argument('--nvrtc-extra', default='',
                    help='extra NVRTC options (tuning), space separated')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
