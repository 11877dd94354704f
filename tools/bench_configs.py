#!/usr/bin/env python
"""Throughput of every BASELINE.json configuration on one GPU (not the
contract bench -- that is bench.py; this fills the table in DESIGN.md).

    python tools/bench_configs.py [--c4-graphs 24] [--c5-graphs 4000]

C1  100 unlabeled graphs, symmetric normalized Gram (all 5 050 pairs)
C2  2000 molecules, symmetric normalized Gram
C3  C2 + Jacobian
C4  NWS graphs of 200-500 nodes with a Convolution node kernel (subset)
C5  X x Y off-diagonal block of C2-style molecules (subset of the 20k set)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.kernel.fix import Normalization  # noqa: E402
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend  # noqa: E402
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel  # noqa: E402


def timed(fn, repeat=2):
    fn()
    best = 1e99
    for _ in range(repeat):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return out, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--c4-graphs', type=int, default=24)
    ap.add_argument('--c5-graphs', type=int, default=4000)
    ap.add_argument('--only', default='')
    ap.add_argument('--c4-grad', action='store_true')
    ap.add_argument('--c4-order', default='natural',
                    choices=['natural', 'rcm', 'random', 'random+rcm',
                             'random+pbr'],
                    help='node order of the C4 graphs (sensitivity of the '
                         'large-pair kernel to the tile structure)')
    args = ap.parse_args()
    be = B200Backend()
    rows = []

    def record(name, pairs, secs, **extra):
        last = be.last
        row = dict(config=name, pairs=pairs, seconds=secs,
                   pairs_per_s=pairs / secs,
                   kernel_ms=last.get('kernel_ms'),
                   small_kernel=last.get('small_kernel'),
                   kernel=last.get('kernel'), grid=last.get('grid'),
                   smem_bytes=last.get('smem_bytes'),
                   cg_iterations_per_pair=last.get('cg_iterations', 0) / max(1, last.get('n_jobs', 1)),
                   **extra)
        rows.append(row)
        print(json.dumps(row), flush=True)

    only = set(args.only.split(',')) if args.only else None

    if not only or 'C1' in only:
        G = make_config_graphs('C1')
        k = Normalization(make_config_kernel('C1', backend=be))
        K, t = timed(lambda: k(G))
        record('C1', len(G) * (len(G) + 1) // 2, t,
               max_abs_dev_from_one=float(np.abs(K - 1).max()))
    if not only or 'C2' in only:
        G = make_config_graphs('C2')
        k = Normalization(make_config_kernel('C2', backend=be))
        K, t = timed(lambda: k(G))
        record('C2', len(G) * (len(G) + 1) // 2, t)
    if not only or 'C3' in only:
        G = make_config_graphs('C2')
        k = Normalization(make_config_kernel('C3', backend=be))
        (K, dK), t = timed(lambda: k(G, eval_gradient=True))
        record('C3', len(G) * (len(G) + 1) // 2, t)
    if not only or 'C4' in only:
        G = make_config_graphs('C4', args.c4_graphs)
        if args.c4_order != 'natural':
            from graphdot_b200.reorder import octile_count, pbr, rcm
            rng = np.random.default_rng(0)
            before = sum(octile_count(g) for g in G)
            if args.c4_order == 'rcm':
                G = [g.permute(rcm(g)) for g in G]
            else:
                G = [g.permute(rng.permutation(len(g.nodes))) for g in G]
                if args.c4_order.endswith('+rcm'):
                    G = [g.permute(rcm(g)) for g in G]
                elif args.c4_order.endswith('+pbr'):
                    G = [g.permute(pbr(g)) for g in G]
            print(json.dumps(dict(order=args.c4_order, octiles_before=before,
                                  octiles_after=sum(octile_count(g)
                                                    for g in G))), flush=True)
        k = make_config_kernel('C4', backend=be)
        if args.c4_grad:
            (K, dK), t = timed(lambda: k(G, eval_gradient=True), repeat=1)
        else:
            K, t = timed(lambda: k(G), repeat=1)
        n = np.array([len(g.nodes) for g in G])
        # SURVEY 8(d) byte model of the large-pair regime: per pair
        # it (40 N + (nnz1 + nnz2) |edge_t|) + 12 N; sum(it N) and sum(it nnz1 nnz2)
        # come from the engine's counters, the edge term is < 1 % and dropped
        iu = np.triu_indices(len(G))
        sum_N = float(np.outer(n, n)[iu].sum())
        model_bytes = 40.0 * be.last['vector_elements'] + 12.0 * sum_N
        gbps = model_bytes / (be.last['kernel_ms'] * 1e-3) / 1e9
        record('C4' + ('+grad' if args.c4_grad else ''),
               len(G) * (len(G) + 1) // 2, t, n_graphs=len(G),
               order=args.c4_order,
               mean_N=float(np.mean(np.outer(n, n))),
               model_GBps=gbps, hbm_peak_GBps=6551.0,
               hbm_frac=gbps / 6551.0,
               products_per_s=be.last['matvec_products']
               / (be.last['kernel_ms'] * 1e-3))
    if only and 'C4ref' in only:
        # the reference's own device code on the same C4 graphs (oracle/_ref/c4_gram)
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import ref_device
        ref = ref_device.RefDeviceSolver('c4_gram')
        n = min(args.c4_graphs, ref.n_graphs)
        jobs = ref_device.triu_jobs(n)
        ref.solve(ref_device.triu_jobs(4), 0.05, n=n)
        Kr, _, ms = ref.solve(jobs, 0.05, n=n)
        G = make_config_graphs('C4', n)
        K = make_config_kernel('C4', backend=be)(G)
        row = dict(config='C4 reference device code', pairs=len(jobs),
                   kernel_ms=ms, pairs_per_s=len(jobs) / (ms * 1e-3),
                   ours_kernel_ms=be.last['kernel_ms'],
                   ours_pairs_per_s=len(jobs) / (be.last['kernel_ms'] * 1e-3),
                   max_rel_diff=float(np.abs(K / Kr - 1).max()), n_graphs=n)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if not only or 'C5' in only:
        G = make_config_graphs('C5', args.c5_graphs)
        h = len(G) // 2
        k = Normalization(make_config_kernel('C5', backend=be))
        K, t = timed(lambda: k(G[:h], G[h:]), repeat=1)
        record('C5', h * (len(G) - h), t, block=f'{h}x{len(G) - h}')
    return rows


if __name__ == '__main__':
    main()
