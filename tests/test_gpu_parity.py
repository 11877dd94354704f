"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the
reference's golden vectors.  Mirrors the reference's hot-path integration
tests (reference test/kernel/marginalized/test_kernel.py:173-605).

Tolerances (BASELINE.json north_star): Gram entries 1e-5 relative, gradients
1e-4 relative, FP32 arithmetic on the device vs the float64 oracle."""
import copy

import numpy as np
import pytest

from conftest import golden_graphs, golden_kernels
from graphdot_b200 import Graph
from graphdot_b200.kernel.fix import Normalization
from graphdot_b200.kernel.marginalized import MarginalizedGraphKernel
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend
from graphdot_b200.kernel.marginalized.starting_probability import Uniform
from graphdot_b200.microkernel import (Constant, KroneckerDelta,
                                       SquareExponential, TensorProduct)
from graphdot_b200.synthetic import make_config_graphs, make_config_kernel
from oracle import mlgk_oracle as oracle

pytestmark = pytest.mark.gpu

CASES = ['unlabeled', 'labeled', 'weighted', 'vario-features', 'molecular',
         'molecular-multitile']
GRAM_RTOL = 1e-5
GRAD_RTOL = 1e-4


@pytest.fixture(scope='module')
def backend():
    return B200Backend()


def rel_err(got, want):
    want = np.asarray(want, float)
    return np.abs(np.asarray(got, float) - want).max() / np.abs(want).max()


@pytest.mark.parametrize('name', CASES)
def test_self_similarity_vs_reference_golden(mlgk_golden, backend, name):
    """reference test_kernel.py:191-217"""
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    for e in case['entries']:
        mlgk = MarginalizedGraphKernel(knode, kedge, q=e['q'],
                                       backend=backend)
        R = mlgk(G)
        assert R.shape == (2, 2)
        assert np.count_nonzero(R - R.T) == 0
        # golden values carry the reference's CG tolerance (rtol 1e-5)
        assert R[0, 0] == pytest.approx(e['K00'], rel=2e-5)
        assert R[1, 1] == pytest.approx(e['K11'], rel=2e-5)
        assert R[0, 1] == pytest.approx(e['K01'], rel=2e-5)
        want = oracle.gram(G, knode=knode, kedge=kedge, q=e['q'])
        assert rel_err(R, want) < GRAM_RTOL
        assert np.allclose(R, want, rtol=GRAM_RTOL)
        d = np.diag(R) ** -0.5
        K = np.diag(d) @ R @ np.diag(d)
        assert K[0, 0] == pytest.approx(1, abs=2e-7)
        assert K[1, 1] == pytest.approx(1, abs=2e-7)


@pytest.mark.parametrize('name', CASES)
def test_cross_similarity_blocks(mlgk_golden, backend, name):
    """reference test_kernel.py:220-241"""
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    for e in case['entries']:
        mlgk = MarginalizedGraphKernel(knode, kedge, q=e['q'],
                                       backend=backend)
        R = mlgk(G)
        assert np.allclose(mlgk(G[:1], G), R[:1, :], rtol=1e-6)
        assert np.allclose(mlgk(G[1:], G), R[1:, :], rtol=1e-6)
        assert np.allclose(mlgk(G, G[:1]), R[:, :1], rtol=1e-6)
        assert np.allclose(mlgk(G, G[1:]), R[:, 1:], rtol=1e-6)


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('lmin', [0, 1])
def test_gradient_vs_oracle_adjoint(mlgk_golden, backend, name, lmin):
    """Analytic Jacobian against the float64 adjoint oracle (which is itself
    checked against central differences in tests/test_oracle.py)."""
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    for q in (0.05, 0.5):
        p = Uniform(1.5)
        mlgk = MarginalizedGraphKernel(knode, kedge, q=q, p=p,
                                       backend=backend)
        R, dR = mlgk(G, eval_gradient=True, lmin=lmin)
        Ro, dRo = oracle.gram(G, knode=knode, kedge=kedge, q=q, p=p,
                              lmin=lmin, eval_gradient=True)
        assert rel_err(R, Ro) < GRAM_RTOL
        mask = mlgk.active_theta_mask
        assert dR.shape == (2, 2, mask.sum())
        dRo = dRo[:, :, mask]
        for k in range(dR.shape[2]):
            scale = np.abs(dRo[:, :, k]).max()
            if scale < 1e-12:
                assert np.abs(dR[:, :, k]).max() < 1e-6
            else:
                assert np.abs(dR[:, :, k] - dRo[:, :, k]).max() \
                    < GRAD_RTOL * scale, (name, q, k)
        # X-by-Y and diag variants agree with the symmetric one
        R2, dR2 = mlgk(G[:1], G, eval_gradient=True, lmin=lmin)
        assert np.allclose(dR2, dR[:1], rtol=1e-5, atol=1e-6)
        D, dD = mlgk.diag(G, eval_gradient=True, lmin=lmin)
        assert np.allclose(D, np.diag(R), rtol=1e-6)
        assert np.allclose(dD, np.einsum('iik->ik', dR), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('name', ['labeled', 'weighted'])
def test_gradient_vs_finite_differences_of_the_kernel(mlgk_golden, backend,
                                                      name):
    """reference test_kernel.py:244-289 (eps 1e-3 in log-theta, 5 %)."""
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    mlgk = MarginalizedGraphKernel(knode, kedge, q=0.1, backend=backend)
    R, dR = mlgk(G, eval_gradient=True)
    theta = mlgk.theta
    for i in range(len(theta)):
        eps = 1e-3
        t = theta.copy()
        t[i] += eps
        mlgk.theta = t
        Rr = mlgk(G)
        t[i] -= 2 * eps
        mlgk.theta = t
        Rl = mlgk(G)
        mlgk.theta = theta
        dR_dt = (Rr - Rl) / (2 * eps) / np.exp(theta)[i]
        assert np.allclose(dR[:, :, i], dR_dt, rtol=0.05, atol=0.05)


@pytest.mark.parametrize('name', ['labeled', 'weighted', 'vario-features',
                                  'molecular'])
@pytest.mark.parametrize('lmin', [0, 1])
def test_nodal_gradient_vs_oracle(mlgk_golden, backend, name, lmin):
    """Nodal Jacobian (forward sensitivities on the device; the reference uses
    finite differences, reference template.cu:226-418, and only checks them to
    5 %, test_kernel.py:244-289) against central differences of the float64
    oracle, for the full matrix, an X-by-Y block and the nodal diagonal."""
    from graphdot_b200.util import flatten, fold_like
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    q, pv = 0.1, 1.3
    mlgk = MarginalizedGraphKernel(knode, kedge, q=q, p=Uniform(pv),
                                   backend=backend)
    R, dR = mlgk(G, nodal=True, eval_gradient=True, lmin=lmin)
    # round 2: the sensitivity solves run in the small-pair kernel on the cached W
    assert backend.last['small_kernel']
    assert dR.shape == (*R.shape, mlgk.active_theta_mask.sum())
    assert np.count_nonzero(dR - dR.transpose(1, 0, 2)) == 0

    def value(pv, qv, tv, te):
        kn, ke = golden_kernels(name)
        kn.theta = fold_like(tv, kn.theta)
        ke.theta = fold_like(te, ke.theta)
        return oracle.gram(G, knode=kn, kedge=ke, q=qv, p=Uniform(pv),
                           nodal=True, lmin=lmin)

    assert rel_err(R, value(pv, q, list(flatten(knode.theta)),
                            list(flatten(kedge.theta)))) < GRAM_RTOL
    base = [np.array([pv]), np.array([q]),
            np.array(list(flatten(knode.theta)), float),
            np.array(list(flatten(kedge.theta)), float)]
    fd = []
    for blk in range(4):
        for i in range(len(base[blk])):
            h = 1e-6 * max(1.0, abs(base[blk][i]))
            hi = [b.copy() for b in base]
            lo = [b.copy() for b in base]
            hi[blk][i] += h
            lo[blk][i] -= h
            fd.append((value(hi[0][0], hi[1][0], hi[2], hi[3]) -
                       value(lo[0][0], lo[1][0], lo[2], lo[3])) / (2 * h))
    fd = np.stack(fd, axis=2)[:, :, mlgk.active_theta_mask]
    for k in range(dR.shape[2]):
        scale = np.abs(fd[:, :, k]).max()
        assert np.abs(dR[:, :, k] - fd[:, :, k]).max() < \
            GRAD_RTOL * max(scale, 1e-6), (name, lmin, k)
    n0 = len(G[0].nodes)
    Rxy, dRxy = mlgk(G[:1], G[1:], nodal=True, eval_gradient=True, lmin=lmin)
    assert np.allclose(dRxy, dR[:n0, n0:], rtol=1e-4, atol=1e-6)
    D, dD = mlgk.diag(G, nodal=True, eval_gradient=True, lmin=lmin)
    assert np.allclose(D, np.diag(R), rtol=1e-6)
    assert np.allclose(dD, np.einsum('iik->ik', dR), rtol=1e-4, atol=1e-6)


def test_nodal_gradient_small_kernel_matches_general_kernel(monkeypatch):
    """The same nodal Jacobian from the two kernels that implement it: the
    small-pair kernel (sensitivity rounds on the cached W, two parameters per
    round) and the general kernel, on molecular graphs incl. nodal='block'."""
    G = make_config_graphs('C2', 10)
    be = B200Backend()
    k1 = make_config_kernel('C3', backend=be)
    R1, dR1 = k1(G, nodal=True, eval_gradient=True)
    assert be.last['small_kernel']
    B1 = k1.diag(G[:4], nodal='block', eval_gradient=False)
    monkeypatch.setenv('GDB_FORCE_GENERAL', '1')
    be2 = B200Backend()
    k2 = make_config_kernel('C3', backend=be2)
    R2, dR2 = k2(G, nodal=True, eval_gradient=True)
    assert not be2.last['small_kernel']
    assert rel_err(R1, R2) < 2e-6
    for m in range(dR1.shape[2]):
        assert rel_err(dR1[:, :, m], dR2[:, :, m]) < 2e-5, m
    B2 = k2.diag(G[:4], nodal='block', eval_gradient=False)
    assert all(rel_err(a, b) < 2e-6 for a, b in zip(B1, B2))
    X1, dX1 = k1(G[:3], G[3:7], nodal=True, eval_gradient=True, lmin=1)
    X2, dX2 = k2(G[:3], G[3:7], nodal=True, eval_gradient=True, lmin=1)
    assert rel_err(X1, X2) < 2e-6
    for m in range(dX1.shape[2]):
        assert rel_err(dX1[:, :, m], dX2[:, :, m]) < 2e-5, m


@pytest.mark.parametrize('name', CASES)
def test_diag_and_nodal(mlgk_golden, backend, name):
    """reference test_kernel.py:292-340"""
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    for e in case['entries']:
        q = e['q']
        mlgk = MarginalizedGraphKernel(knode, kedge, q=q, backend=backend)
        R = mlgk(G)
        D = mlgk.diag(G)
        assert D == pytest.approx(np.diag(R), rel=1e-7)
        Rn = mlgk(G, nodal=True)
        n = np.array([len(g.nodes) for g in G])
        ends = np.cumsum(n)
        assert Rn.shape == (ends[-1], ends[-1])
        assert np.count_nonzero(Rn - Rn.T) == 0
        want = oracle.gram(G, knode=knode, kedge=kedge, q=q, nodal=True)
        assert rel_err(Rn, want) < GRAM_RTOL
        a, b = 0, ends[0]
        assert rel_err(Rn[a:b, a:b], e['nodal00']) < 2e-5
        assert rel_err(Rn[b:, b:], e['nodal11']) < 2e-5
        assert rel_err(Rn[a:b, b:], e['nodal01']) < 2e-5
        Dn = mlgk.diag(G, nodal=True)
        assert Dn == pytest.approx(np.diag(Rn), rel=1e-7)
        blocks = mlgk.diag(G, nodal='block')
        assert len(blocks) == 2
        for blk, (s, t) in zip(blocks, [(0, ends[0]), (ends[0], ends[1])]):
            assert np.allclose(blk, Rn[s:t, s:t], rtol=1e-6)
        Rxy = mlgk(G[:1], G[1:], nodal=True)
        assert np.allclose(Rxy, Rn[:ends[0], ends[0]:], rtol=1e-6)


@pytest.mark.parametrize('name', CASES[:4])
def test_lmin_identity(mlgk_golden, backend, name):
    """reference test_kernel.py:389-408: R0 = R1 + knode"""
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    for q in (0.01, 0.5):
        mlgk = MarginalizedGraphKernel(knode, kedge, q=q, backend=backend)
        g = G[0]
        R0 = mlgk([g], nodal=True, lmin=0)
        R1 = mlgk([g], nodal=True, lmin=1)
        for i, n1 in g.nodes.iterrows():
            for j, n2 in g.nodes.iterrows():
                assert R0[i, j] == pytest.approx(R1[i, j] + knode(n1, n2),
                                                 abs=1e-5 * abs(R0).max())


@pytest.mark.parametrize('name', CASES[:4])
def test_starting_probability(mlgk_golden, backend, name):
    """reference test_kernel.py:411-439: p = 2 scales K by 4"""
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    e = case['entries'][1]
    mlgk = MarginalizedGraphKernel(knode, kedge, q=e['q'], p=2.0,
                                   backend=backend)
    R = mlgk(G)
    assert R[0, 0] == pytest.approx(4 * e['K00'], rel=2e-5)
    assert R[1, 1] == pytest.approx(4 * e['K11'], rel=2e-5)
    adhoc = MarginalizedGraphKernel(
        knode, kedge, q=e['q'], backend=backend,
        p=(lambda nodes: 2.0 * np.ones(len(nodes)), '2.0f'))
    assert np.allclose(adhoc(G), R, rtol=1e-6)


def test_random_weighted_graphs_with_self_loops(backend):
    """reference test_kernel.py:507-525: dense graphs, negative weights and
    self loops (tolerance 5e-4 as in the reference)."""
    import networkx as nx
    knode, kedge = Constant(1.0), Constant(1.0)
    mlgk = MarginalizedGraphKernel(knode, kedge, q=0.1, backend=backend)
    rng = np.random.RandomState(2)
    for _ in range(10):
        n = rng.randint(4, 20)
        A = rng.randn(n, n)
        A = A + A.T
        G = [Graph.from_networkx(nx.from_numpy_array(A), weight='weight')]
        K = mlgk(G).item()
        K0 = oracle.gram(G, knode=knode, kedge=kedge, q=0.1).item()
        assert K == pytest.approx(K0, rel=5e-4)


def test_24_node_random_graph(backend):
    """reference test_kernel.py:442-462"""
    import networkx as nx
    rng = np.random.RandomState(0)
    g = nx.Graph()
    n = 24
    for i, row in enumerate(rng.randint(0, 2, (n, n))):
        g.add_node(i, type=0)
        for j, pred in enumerate(row[:i]):
            if pred:
                g.add_edge(i, j, weight=1)
    dfg = Graph.from_networkx(g, weight='weight')
    knode = TensorProduct(type=KroneckerDelta(1.0))
    kedge = Constant(1.0)
    for dtype in (float, np.float32, np.float64):
        mlgk = MarginalizedGraphKernel(knode, kedge, q=0.5, dtype=dtype,
                                       backend=backend)
        dot = mlgk([dfg])
        assert dot.shape == (1, 1) and dot.dtype == dtype
        assert mlgk.diag([dfg]).dtype == dtype
        gold = oracle.gram([dfg], knode=knode, kedge=kedge, q=0.5)
        assert dot.item() == pytest.approx(gold.item(), rel=GRAM_RTOL)


def test_permutation_invariance(backend):
    """reference test_kernel.py:492-504 (on a synthetic molecule)."""
    g = make_config_graphs('C2', 1)[0]
    kernel = make_config_kernel('C2', backend=backend)
    rng = np.random.default_rng(0)
    base = kernel([g]).item()
    for _ in range(5):
        h = g.permute(rng.permutation(len(g.nodes)))
        assert kernel([g], [h]).item() == pytest.approx(base, rel=1e-5)


def test_fixed_hyperparameters(backend):
    """reference test_kernel.py:528-569"""
    import networkx as nx
    g = nx.Graph()
    g.add_node(0, feature=0)
    g.add_node(1, feature=1)
    g.add_node(2, feature=0)
    g.add_edge(0, 1, attribute=1.0)
    g.add_edge(0, 2, attribute=2.0)
    G = [Graph.from_networkx(g)]
    kV = TensorProduct(feature=KroneckerDelta(0.5))
    kF = TensorProduct(feature=KroneckerDelta(0.5, h_bounds='fixed'))
    eV = TensorProduct(attribute=SquareExponential(1.0))
    eF = TensorProduct(attribute=SquareExponential(
        1.0, length_scale_bounds='fixed'))
    VV = MarginalizedGraphKernel(kV, eV, backend=backend)
    VF = MarginalizedGraphKernel(kV, eF, backend=backend)
    FF = MarginalizedGraphKernel(kF, eF, backend=backend)
    assert len(VV.theta) == len(VF.theta) + 1 == len(FF.theta) + 2
    Rvv, dvv = VV(G, eval_gradient=True)
    Rvf, dvf = VF(G, eval_gradient=True)
    Rff, dff = FF(G, eval_gradient=True)
    assert Rvv == pytest.approx(Rvf) and Rvv == pytest.approx(Rff)
    assert dvv[:, :, VF.active_theta_mask] == pytest.approx(dvf)
    assert dvv[:, :, FF.active_theta_mask] == pytest.approx(dff)


def test_c1_closed_form_and_normalization(backend):
    """BASELINE config C1 (the reference's CPU-runnable case): unlabeled
    graphs have K = n1 n2 / (1 - (1-q)^2) and a normalized Gram of ones."""
    G = make_config_graphs('C1', 40)
    kernel = make_config_kernel('C1', backend=backend)
    K = kernel(G)
    n = np.array([len(g.nodes) for g in G], float)
    want = np.outer(n, n) / (1 - 0.95 ** 2)
    assert rel_err(K, want) < GRAM_RTOL
    Kn = Normalization(kernel)(G)
    assert np.allclose(Kn, 1.0, atol=2e-6)


def test_c2_sample_vs_oracle_and_bitwise_reproducibility(backend):
    """BASELINE config C2/C3 on a 12-graph sample: Gram and Jacobian against
    the oracle; the gather matvec has no atomics, so a repeated call is
    bit-identical (the reference is not)."""
    G = make_config_graphs('C2', 12)
    kernel = make_config_kernel('C2', backend=backend)
    K, dK = kernel(G, eval_gradient=True)
    Ko, dKo = oracle.gram(G, knode=kernel.node_kernel,
                          kedge=kernel.edge_kernel, q=kernel.q,
                          eval_gradient=True)
    assert np.allclose(K, Ko, rtol=GRAM_RTOL)
    dKo = dKo[:, :, kernel.active_theta_mask]
    for k in range(dK.shape[2]):
        assert np.abs(dK[:, :, k] - dKo[:, :, k]).max() \
            < GRAD_RTOL * np.abs(dKo[:, :, k]).max()
    K2, dK2 = kernel(G, eval_gradient=True)
    assert np.array_equal(K, K2) and np.array_equal(dK, dK2)
    # normalized Gram + gradient (fix.Normalization) against the oracle
    Kn, dKn = Normalization(kernel)(G, eval_gradient=True)
    d = np.sqrt(np.diag(Ko))
    assert np.allclose(Kn, Ko / np.outer(d, d), rtol=GRAM_RTOL)
    assert np.allclose(np.diag(Kn), 1.0, atol=1e-6)
    Kxy, dKxy = Normalization(kernel)(G[:5], G[5:], eval_gradient=True)
    assert np.allclose(Kxy, Kn[:5, 5:], rtol=1e-5)
    assert np.allclose(dKxy, dKn[:5, 5:], rtol=1e-3, atol=1e-5)


def test_c4_large_pair_uses_global_arena(backend):
    """BASELINE config C4 (200-500 nodes, Convolution node kernel): pairs do
    not fit in shared memory; checked on the two smallest graphs of a sample
    against the dense oracle."""
    from graphdot_b200.synthetic import newman_watts_strogatz
    rng = np.random.default_rng(44)
    G = [newman_watts_strogatz(rng, 120), newman_watts_strogatz(rng, 131)]
    kernel = make_config_kernel('C4', backend=backend)
    K = kernel(G)                       # N = 15 720: 5 vectors > 227 KB
    R, K01 = oracle.solve_pair(G[0], G[1], kernel.node_kernel,
                               kernel.edge_kernel, kernel.q)   # sparse LU
    assert K[0, 1] == pytest.approx(K01, rel=GRAM_RTOL)
    assert K[0, 1] == K[1, 0]
    Kn = kernel(G[:1], G[1:], nodal=True)
    assert rel_err(Kn, R) < GRAM_RTOL


def test_implicit_job_grids_match_explicit_lists(backend):
    """Device-side decoding of rectangular / triangular job grids (incl. a
    row-block tile written into a tile-sized output) against explicit (i, j)
    lists, bit for bit."""
    from graphdot_b200.kernel.marginalized._backend_b200 import PairJobs
    from graphdot_b200.util import Timer
    G = make_config_graphs('C2', 9)
    kernel = make_config_kernel('C2', backend=backend)
    K = kernel(G)
    T = MarginalizedGraphKernel.traits
    gs = backend.graphset(G)

    def run(jobs, traits, nX, nY, row0=0, col0=0):
        prog = backend.program(gs, kernel.node_kernel, kernel.edge_kernel,
                               kernel.p, traits)
        out = backend.zeros(nX * nY, np.float32)
        starts = np.arange(len(G) + 1, dtype=np.uint32)
        backend.launch(gs, prog, kernel.node_kernel, kernel.edge_kernel,
                       kernel.p, kernel.q, kernel.eps, kernel.ftol,
                       kernel.gtol, jobs, starts, out, None, nX, nY, 5,
                       row0=row0, col0=col0)
        return out.reshape(nX, nY, order='F').astype(float)

    sym = T(symmetric=True)
    a = run(PairJobs.triu(0, 9), sym, 9, 9)
    b = run(np.asarray(PairJobs.triu(0, 9)), sym, 9, 9)
    assert np.array_equal(a, b) and np.array_equal(a, K)
    tile = run(PairJobs.triu(3, 6, 9), T(), 3, 9, row0=3)
    for i in range(3, 6):
        assert np.array_equal(tile[i - 3, i:], K[i, i:])
        assert np.all(tile[i - 3, :i] == 0)
    rect = run(PairJobs.rect(2, 4, 5, 9), T(), 2, 4, row0=2, col0=5)
    assert np.array_equal(rect, K[2:4, 5:9])


def test_fused_normalization_matches_host_formulas(backend):
    """Normalization fused into the solver epilogue (self-similarities kept
    on the device) against the reference's host formulas (reference
    kernel/fix.py:46-73) applied to the raw outputs, and against the tiled
    multi-worker evaluation."""
    from graphdot_b200.kernel.marginalized._tiles import gram_tiled
    G = make_config_graphs('C2', 10)
    kernel = make_config_kernel('C2', backend=backend)
    norm = Normalization(kernel)
    for X, Y in ((G, None), (G[:4], G[4:])):
        K, dK = norm(X, Y, eval_gradient=True)
        Kh, dKh = norm._host_normalized(X, Y, eval_gradient=True)
        assert np.allclose(K, Kh, rtol=2e-6)
        assert np.allclose(dK, dKh, rtol=1e-4, atol=2e-6)
        assert np.allclose(norm(X, Y), Kh, rtol=2e-6)
    K, dK = norm(G, eval_gradient=True)
    assert np.allclose(np.diag(K), 1.0, atol=1e-6)
    assert np.count_nonzero(K - K.T) == 0
    Kt, dKt = gram_tiled(kernel, G, devices=(0,), eval_gradient=True,
                         tile_rows=4)
    assert np.allclose(Kt, K, rtol=1e-6)
    assert np.allclose(dKt[:, :, kernel.active_theta_mask], dK, rtol=1e-5,
                       atol=1e-6)
    raw = gram_tiled(kernel, G, devices=(0,), normalize=False, tile_rows=3)
    assert np.allclose(raw, kernel(G), rtol=1e-6)
    # nodal outputs are normalized on the host as in the reference
    Kn = norm(G[:2], nodal=True)
    assert Kn.shape[0] == sum(len(g.nodes) for g in G[:2])


def test_block_sizes_agree(backend):
    G = make_config_graphs('C2', 6)
    ref = None
    for block in (32, 64, 128, 256):
        be = B200Backend(block_size=block)
        K = make_config_kernel('C2', backend=be)(G)
        if ref is None:
            ref = K
        assert np.allclose(K, ref, rtol=2e-6)


def test_clone_with_theta_shares_engine(backend):
    kernel = make_config_kernel('C2', backend=backend)
    clone = kernel.clone_with_theta(kernel.theta + 0.1)
    assert clone.backend is backend
    assert copy.deepcopy(kernel).backend is backend
    G = make_config_graphs('C2', 3)
    assert not np.allclose(kernel(G), clone(G))


@pytest.mark.parametrize('cap', ['0', '4096'])
def test_placements_agree(backend, monkeypatch, cap):
    """Vectors in shared memory, graphs-only in shared memory and everything
    in global memory give the same answer (GDB_SMEM_CAP limits the dynamic
    shared memory the launcher may use)."""
    G = make_config_graphs('C2', 5)
    kernel = make_config_kernel('C2', backend=backend)
    K, dK = kernel(G, eval_gradient=True)
    monkeypatch.setenv('GDB_SMEM_CAP', cap)
    K2, dK2 = kernel(G, eval_gradient=True)
    assert np.allclose(K, K2, rtol=1e-6)
    assert np.allclose(dK, dK2, rtol=1e-5, atol=1e-6)


def _path_graph(n, chords=()):
    """Unlabeled path on n nodes (+ optional chords)."""
    edges = [(i, i + 1) for i in range(n - 1)] + list(chords)
    e = np.array(edges, dtype=np.uint32)
    return Graph({'!i': np.arange(n, dtype=np.uint32)},
                 {'!i': e[:, 0], '!j': e[:, 1]}, title=f'path{n}')


def test_ragged_sizes_across_tile_boundaries(backend):
    """Graphs whose sizes straddle the 8-row tile and 32-lane boundaries
    (2 ... 40 nodes) in one call: unlabeled closed form
    K = n1 n2 / (1 - (1-q)^2), every pair, normalized Gram of ones."""
    sizes = [2, 3, 7, 8, 9, 15, 16, 17, 24, 31, 32, 33, 40]
    G = [_path_graph(n, chords=[(0, n - 1)] if n > 3 else ()) for n in sizes]
    for q in (0.05, 0.5):
        kernel = MarginalizedGraphKernel(Constant(1.0), Constant(1.0), q=q,
                                         backend=backend)
        K = kernel(G)
        n = np.array(sizes, float)
        want = np.outer(n, n) / (1 - (1 - q) ** 2)
        assert rel_err(K, want) < GRAM_RTOL
        assert np.allclose(Normalization(kernel)(G), 1.0, atol=2e-6)
        K2, dK2 = kernel(G[:5], G[5:], eval_gradient=True)
        assert rel_err(K2, want[:5, 5:]) < GRAM_RTOL


def test_many_tiny_graphs_job_decoding_at_scale(backend):
    """3000 tiny graphs = 4.5 M pairs: exercises the on-device decoding of the
    triangular / rectangular job grids at large indices; every entry of the
    Gram must be written and equal the closed form."""
    rng = np.random.default_rng(3)
    sizes = rng.integers(2, 5, 3000)
    protos = {n: _path_graph(n) for n in (2, 3, 4)}
    G = [protos[int(n)].copy(deep=True) for n in sizes]
    kernel = MarginalizedGraphKernel(Constant(1.0), Constant(1.0), q=0.3,
                                     backend=backend)
    K = kernel(G)
    want = np.outer(sizes, sizes) / (1 - 0.7 ** 2)
    assert np.allclose(K, want, rtol=GRAM_RTOL)
    Kxy = kernel(G[:1200], G[1200:])
    assert np.allclose(Kxy, want[:1200, 1200:], rtol=GRAM_RTOL)


def test_labeled_ragged_pairs_vs_oracle(backend):
    """Molecular-style labeled graphs of very different sizes against the
    oracle (value and Jacobian)."""
    from graphdot_b200.synthetic import random_molecule
    rng = np.random.default_rng(11)
    G = [random_molecule(rng, n) for n in (2, 5, 9, 17, 30)]
    kernel = make_config_kernel('C2', backend=backend, q=0.1)
    K, dK = kernel(G, eval_gradient=True)
    Ko, dKo = oracle.gram(G, knode=kernel.node_kernel,
                          kedge=kernel.edge_kernel, q=0.1, eval_gradient=True)
    assert np.allclose(K, Ko, rtol=GRAM_RTOL)
    dKo = dKo[:, :, kernel.active_theta_mask]
    for k in range(dK.shape[2]):
        assert np.abs(dK[:, :, k] - dKo[:, :, k]).max() \
            < GRAD_RTOL * np.abs(dKo[:, :, k]).max()


@pytest.mark.parametrize('slots', [2, 4])
@pytest.mark.parametrize('p_edge', [0.1, 0.35, 0.8])
def test_helper_lanes_and_overflow_vs_oracle(slots, p_edge):
    """Columns with more neighbours than one lane gathers borrow helper lanes;
    when the lanes run out the owner walks the rest (overflow).  Sparse to
    nearly complete labeled graphs, both slot widths, value and Jacobian."""
    from graphdot_b200.synthetic import random_labeled_graph
    rng = np.random.default_rng(int(100 * p_edge) + slots)
    sizes = {0.1: (3, 8, 13, 21, 24, 30), 0.35: (3, 8, 11, 13, 15, 17),
             0.8: (3, 5, 6, 9, 11, 12)}[p_edge]    # W must fit in shared memory
    G = [random_labeled_graph(rng, n, p_edge) for n in sizes]
    be = B200Backend(slots_per_lane=slots)
    kernel = make_config_kernel('C2', backend=be, q=0.2)
    K, dK = kernel(G, eval_gradient=True)
    assert be.last['small_kernel']
    Ko, dKo = oracle.gram(G, knode=kernel.node_kernel,
                          kedge=kernel.edge_kernel, q=0.2, eval_gradient=True)
    assert np.allclose(K, Ko, rtol=GRAM_RTOL)
    dKo = dKo[:, :, kernel.active_theta_mask]
    for k in range(dK.shape[2]):
        assert np.abs(dK[:, :, k] - dKo[:, :, k]).max() \
            < GRAD_RTOL * np.abs(dKo[:, :, k]).max()
    # without Jacobian (two workers per thread for n > 32 is exercised by C1)
    assert np.allclose(kernel(G), Ko, rtol=GRAM_RTOL)
    Kn = kernel(G[:3], nodal=True)
    Kno = oracle.gram(G[:3], knode=kernel.node_kernel,
                      kedge=kernel.edge_kernel, q=0.2, nodal=True)
    assert np.allclose(Kn, Kno, rtol=GRAM_RTOL, atol=1e-7)


@pytest.mark.parametrize('lmin', [0, 1])
def test_rectangular_blocks_with_role_choice_vs_oracle(backend, lmin):
    """X x Y blocks whose pairs need both role assignments (the smaller graph
    provides the rows): values, Jacobian and starting-probability Jacobian
    against the oracle; K(X, Y) = K(Y, X)^T."""
    from graphdot_b200.synthetic import random_molecule
    rng = np.random.default_rng(23 + lmin)
    X = [random_molecule(rng, n) for n in (24, 6, 17, 9)]
    Y = [random_molecule(rng, n) for n in (8, 23, 16, 3, 12)]
    kernel = make_config_kernel('C2', backend=backend, q=0.1)
    K, dK = kernel(X, Y, eval_gradient=True, lmin=lmin)
    Ko, dKo = oracle.gram(X, Y, knode=kernel.node_kernel,
                          kedge=kernel.edge_kernel, q=0.1, lmin=lmin,
                          eval_gradient=True)
    assert np.allclose(K, Ko, rtol=GRAM_RTOL)
    dKo = dKo[:, :, kernel.active_theta_mask]
    for k in range(dK.shape[2]):
        assert np.abs(dK[:, :, k] - dKo[:, :, k]).max() \
            < GRAD_RTOL * np.abs(dKo[:, :, k]).max()
    Kt, dKt = kernel(Y, X, eval_gradient=True, lmin=lmin)
    assert np.allclose(Kt.T, K, rtol=2e-6)
    assert np.allclose(np.swapaxes(dKt, 0, 1), dK, rtol=2e-5, atol=1e-6)
