"""The drop-in, exercised on hardware: the REFERENCE's own front end --
``graphdot.kernel.marginalized.MarginalizedGraphKernel`` with the reference's
microkernels, starting probabilities and ``Graph`` objects -- runs on
``backend=B200Backend()`` (plug point: reference
graphdot/kernel/marginalized/_backend_factory.py:6-8; call site: reference
_kernel.py:224-242, :363-381), and returns the same Gram matrices, bit for bit, (and the same Jacobians to
2e-5) as this package's front end on the same fixtures (reference
test/kernel/marginalized/test_kernel.py:129-170).

The reference package is imported from the git-ignored ``baseline/_ref``
(staged by tests/golden/stage_reference.py in the build container) under the
numpy-2 / pycuda-free import shim.  Test infrastructure only."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_graphs, golden_kernels
from graphdot_b200.kernel.marginalized import MarginalizedGraphKernel
from graphdot_b200.kernel.marginalized._backend_b200 import B200Backend

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, 'baseline', '_ref')
CASES = ['unlabeled', 'labeled', 'weighted', 'vario-features']


@pytest.fixture(scope='module')
def ref():
    if not os.path.isdir(os.path.join(REF, 'graphdot')):
        pytest.skip('baseline/_ref not staged (tests/golden/'
                    'stage_reference.py needs /root/reference)')
    sys.path.insert(0, GOLDEN)
    import _refshim
    _refshim.install(REF)
    spec = importlib.util.spec_from_file_location(
        'ref_test_kernel',
        os.path.join(REF, 'test/kernel/marginalized/test_kernel.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from graphdot.kernel.marginalized import MarginalizedGraphKernel as RefMGK
    from graphdot.kernel.marginalized._backend import Backend as RefBackend
    from graphdot.kernel.fix import Normalization as RefNormalization
    assert os.path.realpath(sys.modules['graphdot'].__file__).startswith(
        os.path.realpath(REF))
    RefBackend.register(B200Backend)      # isinstance() in backend_factory
    return dict(cases=mod.case_dict, MGK=RefMGK, Norm=RefNormalization)


@pytest.fixture(scope='module')
def backend():
    return B200Backend()


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('q', [0.01, 0.5])
def test_reference_front_end_on_b200_backend(ref, backend, mlgk_golden, name,
                                             q):
    case = ref['cases'][name]
    G_ref = case['graphs']
    ref_kernel = ref['MGK'](case['knode'], case['kedge'], q=q,
                            backend=backend)
    assert ref_kernel.backend is backend
    G = golden_graphs(mlgk_golden['cases'][name])
    knode, kedge = golden_kernels(name)
    own_kernel = MarginalizedGraphKernel(knode, kedge, q=q, backend=backend)

    # Gram matrix, symmetric and X-by-Y
    R = ref_kernel(G_ref)
    assert backend.last['n_jobs'] == 3
    assert np.array_equal(R, own_kernel(G))
    assert np.array_equal(ref_kernel(G_ref[:1], G_ref),
                          own_kernel(G[:1], G))
    # gradient: every hyper-parameter, masked by the reference's own front end
    R2, dR = ref_kernel(G_ref, eval_gradient=True)
    K2, dK = own_kernel(G, eval_gradient=True)
    assert dR.shape == dK.shape
    assert np.array_equal(R2, K2)
    # the two front ends hand over differently spelled (equivalent)
    # expressions -- expf(-0.5F*d2/l^2) vs exp2f(d2*c) -- so planes that live
    # deep in the tail of the exponential (1e-20) differ by the rounding of
    # its argument, a few 1e-6 relative; the north star's gradient tolerance
    # is 1e-4.  Per plane, relative to the plane's largest entry:
    for m in range(dK.shape[2]):
        scale = np.abs(dK[:, :, m]).max()
        assert np.abs(dR[:, :, m] - dK[:, :, m]).max() <= 2e-5 * scale
    # diag, nodal, lmin
    assert np.array_equal(ref_kernel.diag(G_ref), own_kernel.diag(G))
    assert np.array_equal(ref_kernel(G_ref, nodal=True),
                          own_kernel(G, nodal=True))
    assert np.array_equal(ref_kernel(G_ref, lmin=1), own_kernel(G, lmin=1))
    d_ref, dd_ref = ref_kernel.diag(G_ref, eval_gradient=True, nodal=True)
    d_own, dd_own = own_kernel.diag(G, eval_gradient=True, nodal=True)
    assert np.array_equal(d_ref, d_own)
    assert np.allclose(dd_ref, dd_own, rtol=2e-5,
                       atol=2e-5 * np.abs(dd_own).max())
    # the reference's Normalization decorator on top (host formulas,
    # reference kernel/fix.py:21-74): unit diagonal
    Kn = ref['Norm'](ref_kernel)(G_ref)
    assert np.allclose(np.diag(Kn), 1.0, atol=2e-7)
    # ... and the values the reference's own tests pin
    e = [x for x in mlgk_golden['cases'][name]['entries']
         if x['q'] == q][0]
    assert R[0, 0] == pytest.approx(e['K00'], rel=2e-5)
    assert R[1, 1] == pytest.approx(e['K11'], rel=2e-5)
    assert R[0, 1] == pytest.approx(e['K01'], rel=2e-5)


def test_reference_objects_are_cached_like_the_reference_does(ref, backend):
    """reference _backend_cuda.py:111-116: the packed graph lives in
    graph.cookie[backend.uuid]."""
    G_ref = ref['cases']['labeled']['graphs']
    case = ref['cases']['labeled']
    k = ref['MGK'](case['knode'], case['kedge'], q=0.1, backend=backend)
    k(G_ref)
    assert all(backend.uuid in g.cookie for g in G_ref)
