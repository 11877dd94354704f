"""Host evaluation of the microkernels against values produced by the
reference's own microkernels (tests/golden/microkernel_reference.json), plus
the protocol checks of reference test/microkernel/test_microkernel.py."""
import numpy as np
import pytest

from graphdot_b200.microkernel import (  # noqa: F401  (eval namespace)
    Additive, Composite, Constant, Convolution, DotProduct, KroneckerDelta,
    MicroKernel, Normalize, Product, RationalQuadratic, SquareExponential,
    TensorProduct)
from graphdot_b200.util import flatten

inf = np.inf


def _arg(v):
    if isinstance(v, dict):
        return v
    if isinstance(v, list):
        return np.array(v)
    return v


def test_values_and_jacobians_match_reference(microkernel_golden):
    for item in microkernel_golden['items']:
        k = eval(item['expr'])
        assert list(flatten(k.theta)) == pytest.approx(item['theta'])
        for s in item['samples']:
            x, y = _arg(s['x']), _arg(s['y'])
            f, j = k(x, y, jac=True)
            assert f == pytest.approx(s['f'], rel=1e-12, abs=1e-14), item
            if len(s['jac']) != len(item['theta']):
                # reference bug: Add.__call__ adds the two Jacobian ndarrays
                # element-wise instead of concatenating them (reference
                # graphdot/microkernel/_base.py:205-206); its gen_expr (the
                # device path) concatenates, as we do.
                assert len(j) == len(item['theta'])
                continue
            assert np.allclose(np.asarray(j, float).ravel(), s['jac'],
                               rtol=1e-10, atol=1e-13), item
            assert k(x, y) == pytest.approx(s['f'], rel=1e-12, abs=1e-14)


def test_minmax_matches_reference(microkernel_golden):
    for item in microkernel_golden['items']:
        k = eval(item['expr'])
        got = [None if v is None else float(v) for v in k.minmax]
        want = item['minmax']
        for a, b in zip(got, want):
            assert (a is None and b is None) or a == pytest.approx(b)


@pytest.mark.parametrize('k', [
    Constant(0.5), KroneckerDelta(0.3), SquareExponential(1.5),
    RationalQuadratic(1.0, 2.0), SquareExponential(1.0) + 0.1,
    KroneckerDelta(0.5) * SquareExponential(1.0),
    TensorProduct(a=KroneckerDelta(0.3), b=SquareExponential(1.0)),
    Additive(a=KroneckerDelta(0.3), b=SquareExponential(1.0)).normalized,
    Convolution(KroneckerDelta(0.4)), Product(), DotProduct(),
])
def test_protocol(k):
    assert isinstance(k, MicroKernel)
    assert isinstance(k.name, str)
    # repr round trip
    k2 = eval(repr(k))
    assert repr(k2) == repr(k)
    assert list(flatten(k2.theta)) == pytest.approx(list(flatten(k.theta)))
    # hyper-parameter struct mirror
    dt = k.dtype
    assert dt.isalignedstruct or dt.itemsize == 0
    packed = np.array([k.state], dtype=dt)
    flat = np.frombuffer(packed.tobytes(), dtype=np.float32)
    assert list(flat) == pytest.approx(list(flatten(k.theta)))
    # theta setter
    th = list(flatten(k.theta))
    from graphdot_b200.util import fold_like
    k.theta = fold_like([t * 1.5 for t in th], k.theta)
    assert list(flatten(k.theta)) == pytest.approx([t * 1.5 for t in th])
    f, jac = k.gen_expr('x1', 'x2')
    assert isinstance(f, str) and len(jac) == len(th)


def test_jacobian_is_derivative():
    k = TensorProduct(a=KroneckerDelta(0.3),
                      b=SquareExponential(0.7) + 0.2).normalized
    x, y = {'a': 1, 'b': 0.3}, {'a': 2, 'b': 1.1}
    from graphdot_b200.util import fold_like
    f0, j0 = k(x, y, jac=True)
    th = np.array(list(flatten(k.theta)))
    for i in range(len(th)):
        h = 1e-6
        for s, store in ((+1, 'fp'), (-1, 'fm')):
            t = th.copy()
            t[i] += s * h
            k.theta = fold_like(t, k.theta)
            if s > 0:
                fp = k(x, y)
            else:
                fm = k(x, y)
        k.theta = fold_like(th, k.theta)
        assert (fp - fm) / (2 * h) == pytest.approx(j0[i], rel=1e-5, abs=1e-8)


def test_from_sympy_equals_builtin():
    SE = MicroKernel.from_sympy(
        'SE2', 'square exponential', 'exp(-0.5 * (x - y)**2 * ls**-2)',
        ('x', 'y'), ('ls', np.float32, 1e-6, np.inf, 'length scale'))
    a, b = SE(0.9), SquareExponential(0.9)
    fa, ja = a(0.2, 1.3, jac=True)
    fb, jb = b(0.2, 1.3, jac=True)
    assert fa == pytest.approx(fb) and ja[0] == pytest.approx(jb[0])
    expr, jac = a.gen_expr('x1', 'x2', 'k.')
    assert 'k.ls' in expr and 'x1' in expr and len(jac) == 1
    assert a.bounds == ((1e-6, np.inf),)
    with pytest.raises(KeyError):
        SE()


def test_bounds_validation():
    with pytest.raises(ValueError):
        KroneckerDelta(0.5, h_bounds=0.1)
    assert KroneckerDelta(0.5, h_bounds='fixed').bounds == ('fixed',)
    with pytest.raises(ValueError):
        SquareExponential(1.0) ** SquareExponential(1.0)
    with pytest.raises(ValueError):
        Composite('-', a=Constant(1.0))
    assert Normalize(Normalize(Constant(2.0))).name == 'Normalize'


def test_jacobian_plane_selection_is_a_single_pass_equivalent():
    """MarginalizedGraphKernel._active_planes == jacobian[:, :, mask].astype()
    (values, dtype, Fortran layout) for full and partial masks."""
    import numpy as np
    from graphdot_b200.kernel.marginalized import MarginalizedGraphKernel
    rng = np.random.default_rng(0)
    raw = np.asfortranarray(rng.standard_normal((7, 5, 4)).astype(np.float32))
    for mask in ([True] * 4, [True, False, True, True], [False] * 4):
        for dtype in (np.float64, np.float32):
            got = MarginalizedGraphKernel._active_planes(raw, mask, dtype)
            want = raw[:, :, np.asarray(mask)].astype(dtype)
            assert got.dtype == dtype and got.shape == want.shape
            assert np.array_equal(got, want)
            assert got.flags.f_contiguous or got.size == 0
            assert got is not raw
