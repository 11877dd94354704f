"""CPU oracle for the marginalized graph kernel pair solve.  TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product
(``graphdot_b200``) never does and has no CPU fallback.

float64 numpy restatement of the reference algorithm:

* system assembly follows the reference's own dense oracle ``MLGK``
  (reference test/kernel/marginalized/test_kernel.py:20-68: ``Vx`` :27-29,
  symmetric ``Ex`` fill :31-39, degrees/adjacency :41-54, ``Dx``, ``Ax`` and
  ``linsys = diag(Dx/Vx) - Ax*Ex`` :56-58, right-hand side :60-63, reductions
  :65-68), generalised from one graph to a pair with the product-graph index
  ``i = i1*n2 + i2`` as in reference graphdot/experimental/metric/m3.py:52-106;
* the solve is a direct ``numpy.linalg.solve`` (the reference test uses
  scipy CG with atol 1e-7; the device uses Jacobi-PCG, reference
  graphdot/cpp/marginalized_kernel.h:356-461) -- all three converge to the
  same x;
* post-processing follows reference kernel/marginalized/template.cu:134-204
  (``lmin`` shift, graph-level / nodal reductions with starting
  probabilities);
* the analytic gradient uses the adjoint identities of reference
  graphdot/cpp/marginalized_kernel.h:836-989 (order ``[p..., q, node...,
  edge...]``, :40-46) and is cross-checked against central differences in
  tests/test_oracle.py.  For ``lmin=1`` the oracle differentiates the value
  actually returned (the reference's device code does not, SURVEY 8(a)
  quirks).

Parity pin: checked in tests/test_oracle.py against
tests/golden/mlgk_reference.json, produced by running the reference's own
``MLGK`` in the build container (tests/golden/make_golden.py), and against the
closed form K = p^2 n1 n2 / (1-(1-q)^2) for unlabeled graphs.
"""
import numpy as np


def _ordered_rows(df):
    rows = list(df.rows())
    order = np.argsort(np.asarray(df['!i']))
    return [rows[k] for k in order]


def graph_arrays(g):
    """Degree vector, dense weighted adjacency and, for every *directed*
    nonzero (both orientations, self loops once), its end points and the
    undirected edge row it came from."""
    n = len(g.nodes)
    ei = np.asarray(g.edges['!i']).astype(int)
    ej = np.asarray(g.edges['!j']).astype(int)
    w = (np.asarray(g.edges['!w'], dtype=float) if '!w' in g.edges
         else np.ones(len(ei)))
    deg = np.zeros(n)
    adj = np.zeros((n, n))
    src, dst, eid = [], [], []
    for k, (i, j, wk) in enumerate(zip(ei, ej, w)):
        deg[i] += wk
        if i != j:
            deg[j] += wk
        adj[i, j] = wk
        adj[j, i] = wk
        src.append(i), dst.append(j), eid.append(k)
        if i != j:
            src.append(j), dst.append(i), eid.append(k)
    return deg, adj, np.array(src), np.array(dst), np.array(eid)


def pair_system(g1, g2, knode, kedge, q, jac=False, sparse=False):
    """Dense pieces of the product-graph system of a graph pair.

    Returns a dict with ``D`` (N), ``V`` (N), ``W`` (N, N) such that
    ``A = diag(D/V) - W`` and ``b = D``; with ``jac`` also ``dV`` (n_v, N) and
    ``dW`` (list of n_e (N, N) matrices).  ``sparse`` assembles W as scipy
    CSR (for pairs whose dense N x N system would not fit)."""
    n1, n2 = len(g1.nodes), len(g2.nodes)
    N = n1 * n2
    d1, a1, s1, t1, k1 = graph_arrays(g1)
    d2, a2, s2, t2, k2 = graph_arrays(g2)
    nodes1, nodes2 = _ordered_rows(g1.nodes), _ordered_rows(g2.nodes)
    edges1, edges2 = list(g1.edges.rows()), list(g2.edges.rows())

    V = np.empty(N)
    dV = None
    for i1, u in enumerate(nodes1):
        for i2, v in enumerate(nodes2):
            if jac:
                f, j = knode(u, v, True)
                j = np.asarray(j, float).ravel()
                if dV is None:
                    dV = np.zeros((len(j), N))
                dV[:, i1 * n2 + i2] = j
            else:
                f = knode(u, v)
            V[i1 * n2 + i2] = f

    E = np.empty((len(edges1), len(edges2)))
    dE = None
    for a, e1 in enumerate(edges1):
        for b, e2 in enumerate(edges2):
            if jac:
                f, j = kedge(e1, e2, True)
                j = np.asarray(j, float).ravel()
                if dE is None:
                    dE = np.zeros((len(j), *E.shape))
                dE[:, a, b] = j
            else:
                f = kedge(e1, e2)
            E[a, b] = f

    rows = (s1[:, None] * n2 + s2[None, :]).ravel()
    cols = (t1[:, None] * n2 + t2[None, :]).ravel()
    ww = (a1[s1, t1][:, None] * a2[s2, t2][None, :]).ravel()

    def assemble(vals):
        if sparse:
            import scipy.sparse as sp
            return sp.csr_matrix((vals, (rows, cols)), shape=(N, N))
        M = np.zeros((N, N))
        M[rows, cols] = vals
        return M

    W = assemble(ww * E[k1[:, None], k2[None, :]].ravel())
    out = {'n1': n1, 'n2': n2, 'V': V, 'W': W,
           'dox': np.outer(d1, d2).ravel(),
           'D': np.outer(d1, d2).ravel() / (1.0 - q) ** 2}
    if jac:
        out['dV'] = dV if dV is not None else np.zeros((0, N))
        nE = 0 if dE is None else dE.shape[0]
        out['dW'] = [assemble(ww * dE[m][k1[:, None], k2[None, :]].ravel())
                     for m in range(nE)]
    return out


def _unit_start(nodes):
    """Default starting probability p = 1 on every node (reference
    starting_probability.py:80-81): value and d/dp, evaluated here so that the
    oracle does not lean on the product's implementation."""
    return np.ones(len(nodes)), np.ones((1, len(nodes)))


def _start_prob(p, g):
    vals, dvals = p(g.nodes)
    order = np.argsort(np.asarray(g.nodes['!i']))
    vals = np.asarray(vals, float)[order]
    dvals = np.asarray(dvals, float)
    dvals = dvals[:, order] if dvals.size else np.zeros((0, len(vals)))
    return vals, dvals


def solve_pair(g1, g2, knode, kedge, q, p=None, lmin=0,
               eval_gradient=False, sparse=None):
    """Nodal solution matrix R (n1, n2) with starting probabilities applied,
    graph-level K = R.sum(), and with ``eval_gradient`` the Jacobian of K in
    the order [p..., q, node..., edge...]."""
    p = _unit_start if p is None else p
    if sparse is None:
        sparse = len(g1.nodes) * len(g2.nodes) > 3000
    s = pair_system(g1, g2, knode, kedge, q, jac=eval_gradient, sparse=sparse)
    n1, n2, D, V, W = s['n1'], s['n2'], s['D'], s['V'], s['W']
    if sparse and n1 * n2 > 12000:
        # large pairs (config C4: small-world graphs of 200-500 nodes, N up to
        # 250 000): a sparse LU fills in catastrophically, so the system is
        # solved like the reference's own oracle does (scipy CG, reference
        # test_kernel.py:58-63) -- in float64, Jacobi-preconditioned, and
        # iterated to a relative residual of 1e-14
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        A = (sp.diags(D / V) - W).tocsr()
        Minv = spla.LinearOperator(A.shape, matvec=lambda r: r * (V / D))

        def solve(rhs):
            sol, info = spla.cg(A, rhs, rtol=1e-14, atol=0.0, M=Minv,
                                maxiter=20000)
            if info != 0:
                raise RuntimeError(f'oracle CG did not converge ({info})')
            return sol
    elif sparse:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        lu = spla.splu((sp.diags(D / V) - W).tocsc())
        solve = lu.solve
    else:
        A = np.diag(D / V) - W

        def solve(rhs):
            return np.linalg.solve(A, rhs)
    x = solve(D)
    p1, dp1 = _start_prob(p, g1)
    p2, dp2 = _start_prob(p, g2)
    px = np.outer(p1, p2).ravel()
    xs = x - V if lmin == 1 else x
    R = (xs * px).reshape(n1, n2)
    if not eval_gradient:
        return R, R.sum()

    y = solve(px)                         # A is symmetric
    Q = 1.0 / (1.0 - q)
    grad = []
    for m in range(dp1.shape[0]):
        dpx = (np.outer(dp1[m], p2) + np.outer(p1, dp2[m])).ravel()
        grad.append(dpx @ xs)
    grad.append(y @ (2 * Q * D) - y @ ((2 * Q * D / V) * x))
    for m in range(s['dV'].shape[0]):
        g = np.sum(y * x * D / V ** 2 * s['dV'][m])
        if lmin == 1:
            g -= px @ s['dV'][m]
        grad.append(g)
    for dWm in s['dW']:
        grad.append(y @ (dWm @ x))
    return R, R.sum(), np.array(grad)


def gram(X, Y=None, *, knode, kedge, q, p=None, nodal=False, lmin=0,
         eval_gradient=False):
    """Oracle counterpart of ``MarginalizedGraphKernel.__call__``; the
    Jacobian is only provided for ``nodal=False`` and covers ALL
    hyper-parameters."""
    sym = Y is None
    Y = X if sym else Y
    if nodal:
        r0 = np.concatenate([[0], np.cumsum([len(g.nodes) for g in X])])
        c0 = np.concatenate([[0], np.cumsum([len(g.nodes) for g in Y])])
        K = np.zeros((r0[-1], c0[-1]))
    else:
        K = np.zeros((len(X), len(Y)))
    J = None
    for a, g1 in enumerate(X):
        for b, g2 in enumerate(Y):
            if sym and b < a:
                continue
            res = solve_pair(g1, g2, knode, kedge, q, p, lmin, eval_gradient)
            if nodal:
                K[r0[a]:r0[a + 1], c0[b]:c0[b + 1]] = res[0]
                if sym:
                    K[c0[b]:c0[b + 1], r0[a]:r0[a + 1]] = res[0].T
            else:
                K[a, b] = res[1]
                if sym:
                    K[b, a] = res[1]
            if eval_gradient and not nodal:
                if J is None:
                    J = np.zeros((*K.shape, len(res[2])))
                J[a, b] = res[2]
                if sym:
                    J[b, a] = res[2]
    return (K, J) if eval_gradient else K


def diag(X, *, knode, kedge, q, p=None, nodal=False, lmin=0):
    out = []
    for g in X:
        R, K = solve_pair(g, g, knode, kedge, q, p, lmin)
        out.append(np.diag(R) if nodal is True else
                   (R.ravel() if nodal == 'block' else [K]))
    return np.concatenate(out)


def pcg_fp32(D, V, W, tol=1e-8, maxiter=None):
    """float32 emulation of the device Jacobi-PCG (reference
    graphdot/cpp/marginalized_kernel.h:356-461): returns (x, iterations).
    Used to predict iteration counts for the roofline model."""
    f = np.float32
    D, V, W = D.astype(f), V.astype(f), W.astype(f)
    N = len(D)
    diag = D / V
    x = np.zeros(N, f)
    r = D.copy()
    z = r / diag
    pv = z.copy()
    rho = f(r @ z)
    k = 0
    for k in range(maxiter or N):
        if rho == 0:
            break
        Ap = diag * pv - W @ pv
        pAp = f(pv @ Ap)
        if pAp == 0:
            break
        alpha = rho / pAp
        x += alpha * pv
        r -= alpha * Ap
        z = r / diag
        rho_new = f(r @ z)
        if np.sqrt(f(r @ r)) < tol * N:
            k += 1
            break
        pv = z + (rho_new / rho) * pv
        rho = rho_new
    return x, k
