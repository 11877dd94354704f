"""GPR caller: ``graphdot_b200.model.gaussian_process`` against
golden vectors produced by the reference's own ``GaussianProcessRegressor``
(tests/golden/make_gpr_golden.py), plus the reference's self-consistency tests
(reference test/model/gaussian_process/test_gpr.py:10-46, :69-91, :93-114,
:216-262) restated for this package.  The GPU tests check the device-resident
path (Gram + Jacobian never leave the GPU) against the host path."""
import json
import os
import sys

import numpy as np
import pytest

from graphdot_b200.model.gaussian_process import GaussianProcessRegressor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from gpr_cases import RBF, data  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, 'golden', 'gpr_reference.json')))


@pytest.mark.parametrize('case', GOLD['cases'],
                         ids=lambda c: f"s{c['s']}-L{c['L']}-"
                                       f"{c['regularization']}-"
                                       f"{'ny' if c['normalize_y'] else 'raw'}")
def test_matches_reference_gpr(case):
    X, y, y_masked, Z = data()
    gpr = GaussianProcessRegressor(RBF(case['s'], case['L']), alpha=1e-4,
                                   normalize_y=case['normalize_y'],
                                   regularization=case['regularization'])
    gpr.fit(X, y_masked)
    lml, dlml = gpr.log_marginal_likelihood(eval_gradient=True)
    assert lml == pytest.approx(case['lml'], rel=1e-8, abs=1e-8)
    assert np.allclose(dlml, case['dlml'], rtol=1e-6, atol=1e-8)
    sq, dsq = gpr.squared_loocv_error(eval_gradient=True)
    assert sq == pytest.approx(case['sqloocv'], rel=1e-8)
    assert np.allclose(dsq, case['dsqloocv'], rtol=1e-6, atol=1e-9)
    mean, std = gpr.predict(Z, return_std=True)
    assert np.allclose(mean, case['mean'], rtol=1e-8, atol=1e-10)
    assert np.allclose(std, case['std'], rtol=1e-6, atol=1e-8)
    _, cov = gpr.predict(Z, return_cov=True)
    assert np.allclose(cov, case['cov'], rtol=1e-6, atol=1e-8)
    loo, loo_std = gpr.predict_loocv(X, y_masked, return_std=True)
    assert np.allclose(loo, case['loo'], rtol=1e-8, atol=1e-10)
    assert np.allclose(loo_std, case['loo_std'], rtol=1e-8)


def test_singular_gram_uses_the_clamped_pseudoinverse():
    s = GOLD['singular']
    gpr = GaussianProcessRegressor(RBF(1.0, 1.0), alpha=0, beta=1e-8)
    with pytest.warns(UserWarning, match='pseudoinverse'):
        gpr.fit(np.array(s['X']), np.array(s['y']))
    mean = gpr.predict(np.array([0.0, 0.5, 1.0, 2.0]))
    assert np.allclose(mean, s['mean'], rtol=1e-5, atol=1e-7)
    with pytest.warns(UserWarning):
        assert gpr.log_marginal_likelihood() == pytest.approx(s['lml'],
                                                              rel=1e-6)


def test_constant_inputs_predict_the_mean():
    """reference test_gpr.py:10-46"""
    rng = np.random.default_rng(0)
    X, y = np.ones(3), rng.random(3)
    gpr = GaussianProcessRegressor(RBF(1.0, 1.0), alpha=0)
    with pytest.warns(UserWarning):
        gpr.fit(X, y)
    assert gpr.predict(X) == pytest.approx(np.mean(y))


def test_hyperparameter_optimisation_reaches_the_reference_optimum():
    X, y, _, _ = data()
    gpr = GaussianProcessRegressor(RBF(1.0, 1.0), alpha=1e-4, optimizer=True)
    gpr.fit(X, y, tol=1e-8)
    assert gpr.log_marginal_likelihood() == pytest.approx(
        GOLD['optimized']['lml'], rel=1e-5)
    assert np.allclose(gpr.kernel.theta, GOLD['optimized']['theta'],
                       atol=1e-3)


def test_fit_self_consistency_and_untrained_predict():
    """reference test_gpr.py:69-91"""
    X = np.linspace(-1, 1, 5)
    y = np.sin(X * np.pi)
    gpr = GaussianProcessRegressor(RBF(1.0, 0.7), alpha=1e-12)
    with pytest.raises(RuntimeError):
        gpr.predict(X)
    gpr.fit(X, y)
    z, std = gpr.predict(X, return_std=True)
    assert z == pytest.approx(y, 1e-3, 1e-3)
    assert std == pytest.approx(np.zeros_like(y), 1e-3, 1e-3)
    z, cov = gpr.predict(X, return_cov=True)
    assert cov == pytest.approx(np.zeros((5, 5)), 1e-3, 1e-3)


def test_masked_targets_equal_the_filtered_fit():
    """reference test_gpr.py:93-114"""
    rng = np.random.default_rng(1)
    X = np.arange(10.0)
    y = rng.standard_normal(10)
    y[[1, 4, 7]] = None
    gpr = GaussianProcessRegressor(RBF(1.0, 0.8), alpha=1e-12).fit(X, y)
    ok = ~np.isnan(y)
    base = GaussianProcessRegressor(RBF(1.0, 0.8), alpha=1e-12).fit(X[ok],
                                                                    y[ok])
    grid = np.linspace(-1, 10, 50)
    assert np.allclose(gpr.predict(grid), base.predict(grid))


def test_likelihood_gradient_vs_finite_differences():
    """reference test_gpr.py:216-262"""
    X = np.linspace(-1, 1, 6, endpoint=False)
    y = np.sin(X * np.pi)
    eps = 1e-4
    for L in np.logspace(-1, 0, 6):      # (beyond L = 1 the 6 x 6 Gram matrix is numerically singular)
        kernel = RBF(1.0, L)
        gpr = GaussianProcessRegressor(kernel, alpha=1e-10)
        _, dL = gpr.log_marginal_likelihood(X=X, y=y, eval_gradient=True)
        t0 = np.copy(kernel.theta)
        for k in range(2):
            step = np.zeros(2)
            step[k] = eps
            hi = gpr.log_marginal_likelihood(theta=t0 + step, X=X, y=y)
            lo = gpr.log_marginal_likelihood(theta=t0 - step, X=X, y=y)
            assert dL[k] == pytest.approx((hi - lo) / (2 * eps), 1e-3, 1e-3)


def test_unknown_options_raise():
    with pytest.raises(RuntimeError):
        GaussianProcessRegressor(RBF(1, 1), regularization='?')
    X, y, _, _ = data()
    with pytest.raises(RuntimeError):
        GaussianProcessRegressor(RBF(1, 1), optimizer=True).fit(
            X, y, loss='nope')


# ---------------------------------------------------------------------------
# device-resident path on the marginalized graph kernel
# ---------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_gram_equals_host_gram():
    from graphdot_b200.kernel.fix import Normalization
    from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend
    from graphdot_b200.synthetic import make_config_graphs, make_config_kernel
    G = make_config_graphs('C2', 40)
    kernel = make_config_kernel('C2', backend=B200Backend())
    for k in (kernel, Normalization(kernel)):
        K, dK = k(G, eval_gradient=True)
        Kd, dKd = k.device_gram(G, eval_gradient=True)
        assert Kd.is_cuda and dKd.is_cuda
        assert np.array_equal(Kd.cpu().numpy(), K.astype(np.float32))
        assert np.array_equal(dKd.cpu().numpy(), dK.astype(np.float32))
        assert np.array_equal(k.device_gram(G).cpu().numpy(),
                              k(G).astype(np.float32))
        Kx = k.device_gram(G[:7], G[7:19]).cpu().numpy()
        assert np.array_equal(Kx, k(G[:7], G[7:19]).astype(np.float32))


@pytest.mark.gpu
def test_gpr_on_molecules_device_path_vs_host_path():
    """The same GPR with the linear algebra on the GPU (Gram / Jacobian never
    leave the device) and on the host; likelihood gradient against central
    differences of the likelihood."""
    from graphdot_b200.kernel.fix import Normalization
    from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend
    from graphdot_b200.synthetic import make_config_graphs, make_config_kernel
    G = make_config_graphs('C2', 60)
    rng = np.random.default_rng(3)
    # target: fraction of element-1 nodes (the normalized kernel removes the
    # graph size, composition stays visible through KroneckerDelta(element))
    y = np.array([np.mean(np.asarray(g.nodes['element']) == 1)
                  + 0.002 * rng.standard_normal() for g in G])
    kernel = Normalization(make_config_kernel('C2', backend=B200Backend()))
    dev = GaussianProcessRegressor(kernel, alpha=1e-3, normalize_y=True)
    host = GaussianProcessRegressor(kernel, alpha=1e-3, normalize_y=True,
                                    device='cpu')
    dev.fit(G[:50], y[:50])
    host.fit(G[:50], y[:50])
    assert dev.Kinv.is_cuda and not host.Kinv.is_cuda
    vd, gd = dev.log_marginal_likelihood(eval_gradient=True)
    vh, gh = host.log_marginal_likelihood(eval_gradient=True)
    assert vd == pytest.approx(vh, rel=1e-9)
    assert np.allclose(gd, gh, rtol=1e-7, atol=1e-9)
    sd, gsd = dev.squared_loocv_error(eval_gradient=True)
    sh, gsh = host.squared_loocv_error(eval_gradient=True)
    assert sd == pytest.approx(sh, rel=1e-9)
    assert np.allclose(gsd, gsh, rtol=1e-6, atol=1e-9)
    md, sdv = dev.predict(G[50:], return_std=True)
    mh, shv = host.predict(G[50:], return_std=True)
    assert np.allclose(md, mh, rtol=1e-9) and np.allclose(sdv, shv, atol=1e-9)
    # a sensible regressor (the CPU oracle's Gram matrix gives 0.009 vs 0.104)
    assert np.abs(md - y[50:]).mean() \
        < 0.5 * np.abs(y[50:] - y[:50].mean()).mean()
    # gradient of the objective vs central differences (float32 kernel values:
    # a loose tolerance, as in reference test_gpr.py:262)
    t0 = np.array(kernel.theta)
    eps = 1e-2
    for k in range(len(t0)):
        step = np.zeros_like(t0)
        step[k] = eps
        fd = (dev.log_marginal_likelihood(theta=t0 + step)
              - dev.log_marginal_likelihood(theta=t0 - step)) / (2 * eps)
        assert gd[k] == pytest.approx(fd, rel=5e-2, abs=5e-2 * np.abs(gd).max())
