"""Pins the CPU oracle (oracle/mlgk_oracle.py) to the reference: golden
vectors produced by the reference's own dense oracle ``MLGK`` (reference
test/kernel/marginalized/test_kernel.py:20-68) and the known answers of
BASELINE.md section 4."""
import numpy as np
import pytest

from conftest import golden_graphs, golden_kernels
from oracle import mlgk_oracle as oracle

def _close(got, want, tol=1e-5):
    # the reference's CG bounds the residual norm-wise (rtol 1e-5), so compare
    # norm-wise as well
    want = np.asarray(want)
    return np.abs(got - want).max() <= tol * np.abs(want).max()


CASES = ['unlabeled', 'labeled', 'weighted', 'vario-features', 'molecular',
         'molecular-multitile']


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_mlgk(mlgk_golden, name):
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    for e in case['entries']:
        q = e['q']
        # reference values come from scipy CG with its default rtol=1e-5
        # (and atol=1e-7), hence the 2e-5 tolerances
        R00, K00 = oracle.solve_pair(G[0], G[0], knode, kedge, q)
        R11, K11 = oracle.solve_pair(G[1], G[1], knode, kedge, q)
        R01, K01 = oracle.solve_pair(G[0], G[1], knode, kedge, q)
        assert K00 == pytest.approx(e['K00'], rel=2e-5)
        assert K11 == pytest.approx(e['K11'], rel=2e-5)
        assert K01 == pytest.approx(e['K01'], rel=2e-5)
        assert _close(R00, e['nodal00'])
        assert _close(R11, e['nodal11'])
        assert _close(R01, e['nodal01'])


def test_known_answers_baseline_md(mlgk_golden):
    # BASELINE.md section 4 (reference oracle run in the build container)
    table = {
        'unlabeled': [(452.2613065, 452.2613065), (92.30769231, 92.30769231),
                      (47.36842105, 47.36842105), (12, 12)],
        'labeled': [(11.0020214, 100.8375575), (9.376686034, 20.84356588),
                    (7.981421906, 10.85207799), (4.217465432, 2.965198733)],
        'weighted': [(13.87191787, 201.0050251), (11.71550219, 41.02564103),
                     (9.951923172, 21.05263158), (5.504880953, 5.333333333)],
        'vario-features': [(6.840364034, 5.007556666),
                           (6.558041573, 4.741470873),
                           (6.251744842, 4.460038971),
                           (4.883734231, 3.290932416)],
    }
    for name, rows in table.items():
        case = mlgk_golden['cases'][name]
        G = golden_graphs(case)
        knode, kedge = golden_kernels(name)
        for q, (k0, k1) in zip([0.01, 0.05, 0.1, 0.5], rows):
            assert oracle.solve_pair(G[0], G[0], knode, kedge, q)[1] == \
                pytest.approx(k0, rel=2e-7)
            assert oracle.solve_pair(G[1], G[1], knode, kedge, q)[1] == \
                pytest.approx(k1, rel=2e-7)


def test_closed_form_unlabeled(mlgk_golden):
    G = golden_graphs(mlgk_golden['cases']['unlabeled'])
    knode, kedge = golden_kernels('unlabeled')
    from graphdot_b200.kernel.marginalized.starting_probability import Uniform
    for q in (0.01, 0.3):
        for p in (1.0, 2.0):
            R, K = oracle.solve_pair(G[0], G[1], knode, kedge, q, Uniform(p))
            assert K == pytest.approx(p * p * 9 / (1 - (1 - q) ** 2),
                                      rel=1e-12)


@pytest.mark.parametrize('name', ['labeled', 'weighted', 'vario-features',
                                  'molecular'])
@pytest.mark.parametrize('lmin', [0, 1])
def test_adjoint_gradient_vs_central_differences(mlgk_golden, name, lmin):
    from graphdot_b200.kernel.marginalized.starting_probability import Uniform
    from graphdot_b200.util import flatten, fold_like
    case = mlgk_golden['cases'][name]
    G = golden_graphs(case)
    knode, kedge = golden_kernels(name)
    q, p = 0.07, Uniform(1.3)
    _, K, grad = oracle.solve_pair(G[0], G[1], knode, kedge, q, p, lmin,
                                   eval_gradient=True)

    def value(pv, qv, tv, te):
        kn, ke = golden_kernels(name)
        kn.theta = fold_like(tv, kn.theta)
        ke.theta = fold_like(te, ke.theta)
        return oracle.solve_pair(G[0], G[1], kn, ke, qv, Uniform(pv), lmin)[1]

    tv = np.array(list(flatten(knode.theta)), float)
    te = np.array(list(flatten(kedge.theta)), float)
    base = [np.array([1.3]), np.array([q]), tv, te]
    k = 0
    for blk in range(4):
        for i in range(len(base[blk])):
            h = 1e-6 * max(1.0, abs(base[blk][i]))
            args_p = [b.copy() for b in base]
            args_m = [b.copy() for b in base]
            args_p[blk][i] += h
            args_m[blk][i] -= h
            fd = (value(args_p[0][0], args_p[1][0], args_p[2], args_p[3]) -
                  value(args_m[0][0], args_m[1][0], args_m[2], args_m[3])
                  ) / (2 * h)
            assert grad[k] == pytest.approx(fd, rel=2e-6, abs=1e-7), (blk, i)
            k += 1
    assert k == len(grad)


def test_fp32_pcg_emulation_converges(mlgk_golden):
    case = mlgk_golden['cases']['molecular']
    G = golden_graphs(case)
    knode, kedge = golden_kernels('molecular')
    s = oracle.pair_system(G[0], G[1], knode, kedge, 0.05)
    x, iters = oracle.pcg_fp32(s['D'], s['V'], s['W'])
    ref = np.linalg.solve(np.diag(s['D'] / s['V']) - s['W'], s['D'])
    assert iters < len(ref)
    assert np.allclose(x, ref, rtol=1e-5)
