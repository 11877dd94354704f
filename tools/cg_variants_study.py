#!/usr/bin/env python
"""Offline numerics study (CPU, float32 emulation) for the next kernel round:
does a single-reduction CG (Chronopoulos-Gear) or a symmetrically scaled
system reach the reference's stopping criterion in as many iterations and with
the same accuracy as the Jacobi-PCG the kernels run today?

    python tools/cg_variants_study.py [--pairs 40]

TEST / RESEARCH INFRASTRUCTURE: uses the CPU oracle to assemble the pair
systems of the C2 workload; nothing here is on the product path.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from graphdot_b200.microkernel import (KroneckerDelta, SquareExponential,  # noqa: E402
                                       TensorProduct)
from graphdot_b200.synthetic import make_config_graphs  # noqa: E402
from oracle import mlgk_oracle as oracle  # noqa: E402

f = np.float32


def pcg(diag, W, b, tol, N):
    """Jacobi-PCG as in mlgk_small.cuh (two reductions per iteration)."""
    x = np.zeros(N, f)
    r = b.copy()
    z = r / diag
    p = z.copy()
    rho = f(r @ z)
    for k in range(1, N + 1):
        Ap = diag * p - W @ p
        alpha = rho / f(p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = r / diag
        rho_new = f(r @ z)
        if f(r @ r) < (tol * N) ** 2:
            return x, k
        p = z + (rho_new / rho) * p
        rho = rho_new
    return x, N


def cg_chronopoulos_gear(diag, W, b, tol, N):
    """Preconditioned Chronopoulos-Gear CG: ONE fused reduction per iteration
    (gamma = r.u, delta = w.u, r.r), vectors x, r, p, s and u = M^-1 r, w = A u."""
    x = np.zeros(N, f)
    r = b.copy()
    u = r / diag
    w = diag * u - W @ u
    gamma = f(r @ u)
    delta = f(w @ u)
    p = np.zeros(N, f)
    s = np.zeros(N, f)
    alpha = gamma / delta
    beta = f(0)
    for k in range(1, N + 1):
        p = u + beta * p
        s = w + beta * s
        x += alpha * p
        r -= alpha * s
        u = r / diag
        w = diag * u - W @ u
        gamma_new = f(r @ u)
        delta = f(w @ u)
        rr = f(r @ r)          # same fused reduction
        if rr < (tol * N) ** 2:
            return x, k
        beta = gamma_new / gamma
        alpha = gamma_new / (delta - beta * gamma_new / alpha)
        gamma = gamma_new
    return x, N


def cg_scaled(diag, W, b, tol, N):
    """Plain CG on the symmetrically scaled system D^-1/2 A D^-1/2 (unit
    diagonal); the stopping test uses the unscaled residual."""
    sc = (1 / np.sqrt(diag)).astype(f)
    Ws = (sc[:, None] * W * sc[None, :]).astype(f)
    bs = sc * b
    x = np.zeros(N, f)
    r = bs.copy()
    p = r.copy()
    rho = f(r @ r)
    for k in range(1, N + 1):
        Ap = p - Ws @ p
        alpha = rho / f(p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        rho_new = f(r @ r)
        if f((r * r) @ diag) < (tol * N) ** 2:
            return sc * x, k
        p = r + (rho_new / rho) * p
        rho = rho_new
    return sc * x, N


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=40)
    ap.add_argument('--q', type=float, default=0.05)
    args = ap.parse_args()
    G = make_config_graphs('C2', 2 * args.pairs)
    kn = TensorProduct(element=KroneckerDelta(0.5), x=SquareExponential(1.0))
    ke = TensorProduct(length=SquareExponential(0.1))
    rows = []
    for a in range(args.pairs):
        sysm = oracle.pair_system(G[2 * a], G[2 * a + 1], kn, ke, args.q)
        D, V, W = sysm['D'], sysm['V'], sysm['W']
        N = len(D)
        A = np.diag(D / V) - W
        for name, rhs in (('value', D), ('adjoint', np.ones(N))):
            exact = np.linalg.solve(A, rhs)
            out = [name]
            for fn in (pcg, cg_chronopoulos_gear, cg_scaled):
                x, k = fn((D / V).astype(f), W.astype(f), rhs.astype(f), 1e-8, N)
                err = abs(x.sum(dtype=np.float64) - exact.sum()) / abs(exact.sum())
                out += [k, err]
            rows.append(out)
    for name in ('value', 'adjoint'):
        sel = [r for r in rows if r[0] == name]
        print(f'{name:8s}  iterations (mean)  rel. error of sum(x) (max)')
        for j, label in enumerate(('Jacobi-PCG (today)', 'Chronopoulos-Gear',
                                   'scaled plain CG')):
            its = np.mean([r[1 + 2 * j] for r in sel])
            err = np.max([r[2 + 2 * j] for r in sel])
            print(f'  {label:20s} {its:6.2f}   {err:.2e}')


if __name__ == '__main__':
    main()
