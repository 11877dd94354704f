#!/usr/bin/env python
"""Benchmark of the MLGK hot path (BASELINE.json metric: graph-pair MLGK
solves/sec, Gram + gradient).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm

Workload (``config.workload``): BASELINE config C3 = C2 with
``eval_gradient=True`` -- the normalized Gram matrix AND its Jacobian over all
5 hyper-parameters of the synthetic molecular set (KroneckerDelta(element) x
SquareExponential(x) node kernel, SquareExponential(length) edge kernel,
weighted bonds, q = 0.05).  At N = 1 this is the 2000-graph set of
BASELINE.json (2 001 000 graph pairs per step).  At N > 1 the set grows to
2000*sqrt(N) graphs so that every GPU keeps 2 001 000 pairs per step (weak
scaling); the pairs are cut into row-block tiles that the ranks pull from one
dynamic queue (a counter in the torch.distributed store) -- no collective on
the data path.  A "step" is one pass over all pairs.

``value``   pairs/s with graphs resident in HBM and outputs left on the
            device (CUDA events on the launching stream, max over ranks).
``e2e``     pairs/s through the tile worker with HOST buffers: every step
            re-sends the packed graphs (pinned) host->device, copies every
            Gram / Jacobian tile device->host and normalizes it on the host.
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
import math
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


def n_graphs_for(world):
    return int(round(2000 * math.sqrt(world)))


# ---------------------------------------------------------------------------
# CPU arm (oracle port), also the cpu_baseline of the GPU arm
# ---------------------------------------------------------------------------
def _cpu_pair(args):
    from graphdot_b200.synthetic import make_config_kernel
    from oracle import mlgk_oracle as oracle
    g1, g2 = args
    k = _cpu_pair.kernel
    if k is None:
        k = _cpu_pair.kernel = make_config_kernel('C3', backend=_NoBackend())
    return oracle.solve_pair(g1, g2, k.node_kernel, k.edge_kernel, k.q, k.p,
                             eval_gradient=True)[1]


_cpu_pair.kernel = None


class _NoBackend:
    """Placeholder so the CPU arm can build the kernel object (for its
    hyper-parameters) without touching the CUDA library."""
    def __new__(cls):
        from graphdot_b200.kernel.marginalized._backend import Backend

        class Dummy(Backend):
            def __call__(self, *a, **k):
                raise RuntimeError('CPU arm has no engine')
        return Dummy()


def cpu_sample_rate(n_sample_graphs=16, repeats=1, pool=None):
    """pairs/s of the oracle on the upper-triangle pairs of the first
    ``n_sample_graphs`` C2 graphs, on all host cores."""
    import multiprocessing as mp
    from graphdot_b200.synthetic import make_config_graphs
    G = make_config_graphs('C2', n_sample_graphs)
    pairs = [(G[i], G[j]) for i in range(len(G)) for j in range(i, len(G))]
    cores = os.cpu_count() or 1
    own = pool is None
    if own:
        pool = mp.get_context('fork').Pool(cores)
    try:
        pool.map(_cpu_pair, pairs[:cores])           # warm the workers
        t0 = time.perf_counter()
        for _ in range(repeats):
            pool.map(_cpu_pair, pairs, chunksize=max(1, len(pairs) // (8 * cores)))
        dt = (time.perf_counter() - t0) / repeats
    finally:
        if own:
            pool.close()
            pool.join()
    return len(pairs) / dt, cores, len(pairs), dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    pool = mp.get_context('fork').Pool(cores)
    n_sample = 96
    times = []
    try:
        for s in range(args.warmup + args.steps):
            rate, cores, n_pairs, dt = cpu_sample_rate(n_sample, 1, pool)
            if s >= args.warmup:
                times.append(dt)
    finally:
        pool.close()
        pool.join()
    dt = float(np.mean(times))
    value = n_pairs / dt
    sample = (f'{n_pairs} pairs per step: upper triangle of the first '
              f'{n_sample} graphs of the C2 set, float64 dense Kronecker '
              'solve + adjoint Jacobian')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(world),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def workload_config(world):
    n = n_graphs_for(world)
    return {
        'workload': 'C3 = C2 + eval_gradient: synthetic molecular graphs '
                    '(16-24 nodes), normalized Gram + Jacobian over 5 '
                    'hyper-parameters',
        'n_graphs': n, 'pairs_per_step': n * (n + 1) // 2,
        'node_kernel': 'TensorProduct(element=KroneckerDelta(0.5), '
                       'x=SquareExponential(1.0))',
        'edge_kernel': 'TensorProduct(length=SquareExponential(0.1))',
        'q': 0.05, 'l2': 'flushed between timed steps (256 MiB write)',
        'sharding': 'row-block tiles from a dynamic queue, replicated graphs',
    }


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


def algorithmic_flops(stats, sum_N, sum_nnzx, nJ):
    """FP32 flops of the solver per SURVEY.md 8(d): matvec
    W_mv = nnzx (2 + F_e) + 2N per application, plus 13 N per CG iteration,
    setup N (F_v + 4) per solve, and the Jacobian sweep."""
    it_nnzx = stats['matvec_products']
    it_N = stats['vector_elements']
    matvec = it_nnzx * (2 + F_E) + 2 * it_N
    cg = matvec + 13 * it_N
    setup = sum_N * (F_V + 4) * 2
    jac = sum_nnzx * (2 + F_E + F_DE) + sum_N * (F_DV + 8 + 2 * nJ)
    return {'matvec': matvec, 'total': cg + setup + jac}


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend
    from graphdot_b200.kernel.marginalized._tiles import (GramTileWorker,
                                                          LocalTileQueue,
                                                          StoreTileQueue,
                                                          row_tiles,
                                                          tile_pairs)
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

    n = n_graphs_for(world)
    G = make_config_graphs('C2', n)
    backend = B200Backend(device=local_rank,
                          block_size=args.block_size or None,
                          slots_per_lane=args.slots_per_lane or None,
                          nvrtc_extra=args.nvrtc_extra.split())
    kernel = make_config_kernel('C3', backend=backend)
    stream = torch.cuda.current_stream()
    # row-block height: about 8 tiles per rank -- tall tiles amortise the tail
    # of a launch (measured at N = 1: 32 / 64 / 128 / 256 rows = 20.2 / 20.9 /
    # 21.5 / 21.7 M pairs/s), enough tiles keep the dynamic queue balanced
    tile_rows = args.tile_rows or max(32, min(256, n // (8 * world)))
    worker = GramTileWorker(kernel, G, backend, eval_gradient=True,
                            max_rows=tile_rows,
                            stream=stream.cuda_stream)
    worker.diag(store=True)     # self-similarities for the fused normalization
    tiles = row_tiles(n, tile_rows)
    total_pairs = n * (n + 1) // 2
    assert sum(tile_pairs(a, b, n) for a, b in tiles) == total_pairs
    sizes = np.array([len(g.nodes) for g in G], float)
    nnz = np.array([2 * len(g.edges) for g in G], float)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    step_no = [0]

    def tile_queue():
        """Dynamic tile queue: a local counter (N=1) or an atomic counter in
        the torch.distributed store shared by all ranks (N>1)."""
        key = f'tiles{step_no[0]}'
        step_no[0] += 1
        if store is None:
            return iter(LocalTileQueue(tiles))
        return iter(StoreTileQueue(store, tiles, key))

    def step_device():
        for i0, i1 in tile_queue():
            worker.run_tile(i0, i1, keep_on_device=True, normalize=True)

    def step_e2e():
        # H2D of the packed graphs; self-similarities stay on the device
        worker.diag(upload=True, store=True)
        checksum = 0.0
        for i0, i1 in tile_queue():
            # normalized in the solver epilogue; D2H of the tile
            Kn, dKn = worker.run_tile(i0, i1, normalize=True)
            checksum += float(Kn[0, i0]) + float(dKn[0, i0, 1])
        return checksum

    # ---- device-resident throughput -----------------------------------------
    for _ in range(args.warmup):
        barrier()
        step_device()
    sampler = ClockSampler(local_rank)
    for k in worker.stats:
        worker.stats[k] = 0
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
    stats = dict(worker.stats)

    # ---- end to end with host buffers ----------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        barrier()
        step_e2e()
    for k in worker.stats:
        worker.stats[k] = 0
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_stats = dict(worker.stats)

    # ---- reduce over ranks ------------------------------------------------------
    vec = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device='cuda')
    sums = torch.tensor([stats['kernel_ms'], stats['cg_iterations'],
                         stats['matvec_products'], stats['vector_elements'],
                         stats['launches'], stats['pairs'],
                         e2e_stats['h2d_bytes'], e2e_stats['d2h_bytes']],
                        dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s = vec.tolist()
    (kernel_ms, cg_it, products, vec_elems, launches, pairs_done, h2d,
     d2h) = sums.tolist()
    assert int(pairs_done) == total_pairs * args.steps, (pairs_done,
                                                         total_pairs)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = total_pairs * args.steps / (dev_ms * 1e-3)
    e2e_value = total_pairs * e2e_steps / e2e_s
    # sums over the unique pairs (i <= j) of N = n_i n_j and nnz_i nnz_j
    sum_N = (sizes.sum() ** 2 + (sizes ** 2).sum()) / 2
    sum_nnzx = (nnz.sum() ** 2 + (nnz ** 2).sum()) / 2
    fl = algorithmic_flops({'matvec_products': products / args.steps,
                            'vector_elements': vec_elems / args.steps},
                           sum_N, sum_nnzx, kernel.n_dims)
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
    pinfo = backend.program_info(worker.prog)
    traffic = None      # DRAM bytes of one launch from the committed ncu capture
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')))
        traffic = tr['dram_bytes_per_launch']
    except (OSError, KeyError, ValueError):
        pass
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': dict(workload_config(world), tile_rows=tile_rows,
                       tiles_per_step=len(tiles)),
        'e2e': {'value': e2e_value, 'unit': UNIT,
                'h2d_bytes_per_step': int(h2d / e2e_steps),
                'd2h_bytes_per_step': int(d2h / e2e_steps),
                'steps': e2e_steps,
                'note': 'tile worker with host buffers: graphs H2D, diagonal '
                        'solve, normalized Gram+Jacobian tiles D2H'},
        'gpu_launches': int(launches),
        'roofline': {
            'bound': 'fp32', 'achieved': achieved, 'peak': peak_tflops,
            'unit': 'TFLOP/s', 'frac': achieved / peak_tflops,
            'traffic': traffic,
            'traffic_note': 'dram read+write bytes of ONE launch (row-block tile '
                            'of 125 984 pairs) from profiles/r1_small_final_ncu_'
                            'summary.md; the kernel is on-chip bound, the blobs '
                            'stream in once',
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
            'kernel': (f'mlgk_solve_small block={pinfo.block_size} '
                       f'regs={pinfo.num_regs_small}'
                       if backend.last.get('small_kernel') else
                       f'mlgk_solve block={pinfo.block_size} '
                       f'regs={pinfo.num_regs}'),
        },
        'clocks': clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, n_pairs, dt = cpu_sample_rate(96)
        line['cpu_baseline'] = {
            'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{n_pairs} pairs (upper triangle of the first 96 '
                      f'graphs), {dt:.1f} s, float64 dense Kronecker solve + '
                      'adjoint Jacobian (oracle/mlgk_oracle.py)'}
    if world == 1 and not args.no_reference_gpu:
        line['reference_gpu'] = reference_gpu_rate(n, kernel.q)
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
    ap.add_argument('--tile-rows', type=int, default=0,
                    help='rows per tile launch; 0 = about 8 tiles per rank')
    ap.add_argument('--block-size', type=int, default=0)
    ap.add_argument('--slots-per-lane', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=2)
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
