#!/usr/bin/env python
"""Decision study for the separable tensor-core path (BASELINE north_star:
"Tensor cores are used only for the separable case, where the matvec collapses
to A.R.A'^T dense contractions.  There, TF32/FP32-emulated accuracy must be
shown to stay within tolerance, or the path is dropped").

For Constant / label-free edge kernels (BASELINE config C1, reference
example/unlabeled-unweighted.py, reference microkernel/_base.py:333-385) the
product-graph matvec is  y = (Dx/Vx) o x - c_e A1 X A2^T  with X the n1 x n2
matrix form of x.  This script runs the engine's Jacobi-PCG in float32 numpy
arithmetic with that contraction evaluated

  fp32      in float32 (what the CUDA kernels do),
  tf32x1    as two tensor-core MMAs with TF32 operands (10-bit mantissa,
            round-to-nearest as cvt.rna.tf32.f32 does), FP32 accumulation,
  tf32x3    with the 3xTF32 split (operand = hi + lo, both TF32; A is exactly
            representable for unweighted graphs) -- 2 MMAs per contraction,
            4 per matvec,

and reports the Gram error against the float64 direct solve.  Tolerance of
the north star: 1e-5 relative.  Pure numpy, no GPU.

    python tools/tf32_separable_study.py [--pairs 300]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphdot_b200.synthetic import make_config_graphs  # noqa: E402


def tf32(a):
    """Round float32 to TF32 (10 explicit mantissa bits), nearest-even-ish
    (round half away from zero on the magnitude, like cvt.rna)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    bits = a.view(np.uint32).astype(np.uint64)
    bits = (bits + 0x1000) & 0xFFFFE000
    return bits.astype(np.uint32).view(np.float32)


def adjacency(g):
    n = len(g.nodes)
    A = np.zeros((n, n), np.float32)
    i = np.asarray(g.edges['!i']).astype(int)
    j = np.asarray(g.edges['!j']).astype(int)
    w = (np.asarray(g.edges['!w'], np.float32) if '!w' in g.edges
         else np.ones(len(i), np.float32))
    A[i, j] = w
    A[j, i] = w
    return A


def contraction(A1, X, A2, mode):
    f = np.float32
    if mode == 'fp32':
        return (A1 @ X) @ A2.T
    if mode == 'tf32x1':
        T = (tf32(A1).astype(f) @ tf32(X).astype(f)).astype(f)
        return (tf32(T).astype(f) @ tf32(A2).T.astype(f)).astype(f)
    if mode == 'tf32x3':
        def split(M):
            hi = tf32(M)
            return hi, tf32(M - hi)
        a1h, a1l = split(A1)
        a2h, a2l = split(A2)
        xh, xl = split(X)
        T = (a1h @ xh + a1h @ xl + a1l @ xh).astype(f)
        th, tl = split(T)
        return (th @ a2h.T + tl @ a2h.T + th @ a2l.T).astype(f)
    raise KeyError(mode)


def pcg(A1, A2, q, mode, cv=1.0, ce=1.0, ftol=1e-8):
    """The engine's Jacobi-PCG (graphdot_b200/csrc/mlgk_solver.cuh gdb_pcg),
    float32, separable matvec; returns K = sum x (p = 1) and iterations."""
    f = np.float32
    d1, d2 = A1.sum(1), A2.sum(1)
    D = (np.outer(d1, d2) / f((1 - q) ** 2)).astype(f)
    diag = (D / f(cv)).astype(f)
    x = np.zeros_like(D)
    r = D.copy()
    z = (r / diag).astype(f)
    p = z.copy()
    rho = f((r * z).sum(dtype=f))
    N = D.size
    it = 0
    while it < N and rho != 0:
        Ap = (diag * p - f(ce) * contraction(A1, p, A2, mode)).astype(f)
        pAp = f((p * Ap).sum(dtype=f))
        if pAp == 0:
            break
        it += 1
        alpha = f(rho / pAp)
        x = (x + alpha * p).astype(f)
        r = (r - alpha * Ap).astype(f)
        rr = f((r * r).sum(dtype=f))
        z = (r / diag).astype(f)
        rz = f((r * z).sum(dtype=f))
        if np.sqrt(rr) < ftol * N:
            break
        p = (z + f(rz / rho) * p).astype(f)
        rho = rz
    return float(x.sum(dtype=np.float64)), it


def exact(A1, A2, q, cv=1.0, ce=1.0):
    A1, A2 = A1.astype(float), A2.astype(float)
    D = np.outer(A1.sum(1), A2.sum(1)).ravel() / (1 - q) ** 2
    M = np.diag(D / cv) - ce * np.kron(A1, A2)
    return float(np.linalg.solve(M, D).sum())


def study(n_pairs=300, q=0.05, seed=0, weighted=False):
    G = make_config_graphs('C1')
    rng = np.random.default_rng(seed)
    A = [adjacency(g) for g in G]
    if weighted:      # non-representable weights: the operand A is rounded too
        for M in A:
            W = rng.uniform(0.5, 1.0, M.shape).astype(np.float32)
            W = np.triu(W, 1)
            M *= (W + W.T)
    pairs = [(int(a), int(b)) for a, b in
             zip(rng.integers(0, len(G), n_pairs),
                 rng.integers(0, len(G), n_pairs))]
    out = {}
    ref = [exact(A[a], A[b], q) for a, b in pairs]
    for mode in ('fp32', 'tf32x1', 'tf32x3'):
        res = [pcg(A[a], A[b], q, mode) for a, b in pairs]
        err = [abs(k - e) / abs(e) for (k, _), e in zip(res, ref)]
        out[mode] = dict(max_rel_err=float(np.max(err)),
                         median_rel_err=float(np.median(err)),
                         mean_iterations=float(np.mean([it for _, it in res])),
                         within_1e5=bool(np.max(err) <= 1e-5))
    return dict(config='C1' + ('+weights' if weighted else ''), q=q,
                pairs=len(pairs), **out)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=300)
    args = ap.parse_args()
    for weighted in (False, True):
        for q in (0.05, 0.01):
            print(json.dumps(study(args.pairs, q=q, weighted=weighted)),
                  flush=True)
