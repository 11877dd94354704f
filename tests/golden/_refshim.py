"""Import shim that lets the *reference's own Python* run in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``make_golden.py`` (and nothing else) to
import ``/root/reference/graphdot`` without pycuda/ase/mendeleev and under
numpy >= 2, so that golden vectors come from the reference's own code, and by
``tests/test_gpu_reference_frontend.py`` to import the copy staged under
``baseline/_ref`` (``stage_reference.py``) on the GPU box:

* numpy 2 removed ``np.float/np.int/np.object/np.issctype/np.issubsctype``
  (used at reference graphdot/kernel/marginalized/_kernel.py:62,
  graphdot/minipandas/series.py:12, graphdot/minipandas/dataframe.py:22 ...).
* ``mendeleev`` (graphdot/graph/adjacency/atomic.py:7), ``pymatgen``, ``ase``
  and ``pycuda`` (graphdot/cuda/__init__.py:3) are not installed.

Nothing under /root/reference is copied; it is imported in place.  This module
cannot run on the GPU box (no /root/reference there) and is never imported by
the product or by tests.
"""
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'


def _is_scalar_type(t):
    try:
        dt = np.dtype(t)
    except TypeError:
        return False
    return dt.kind != 'O' and dt.names is None


def install(root=None):
    root = root or REFERENCE_ROOT
    # -- numpy >= 2 aliases ------------------------------------------------
    for name, val in (('float', float), ('int', int), ('object', object),
                      ('bool', bool)):
        if not hasattr(np, name):
            setattr(np, name, val)
    if not hasattr(np, 'issctype'):
        np.issctype = _is_scalar_type
    if not hasattr(np, 'issubsctype'):
        np.issubsctype = lambda a, b: np.issubdtype(np.dtype(a), b)

    # -- absent third-party modules ---------------------------------------
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, k):
            return _Anything()

    stub('mendeleev')
    stub('mendeleev.fetch', fetch_table=lambda *a, **k: _Anything())
    stub('pymatgen')
    stub('pymatgen.io')
    stub('pymatgen.io.ase', AseAtomsAdaptor=_Anything)
    stub('ase')
    stub('ase.build', molecule=_Anything())

    class _Managed(np.ndarray):
        pass

    def managed_empty(shape, dtype, order='C', mem_flags=0):
        return np.empty(shape, dtype, order)

    def managed_zeros(shape, dtype, order='C', mem_flags=0):
        return np.zeros(shape, dtype, order)

    def managed_empty_like(a, mem_flags=0):
        return np.empty_like(a)

    class _Ctx:
        def get_device(self):
            return _Anything()

        def synchronize(self):
            pass

    drv = stub('pycuda.driver',
               managed_empty=managed_empty, managed_zeros=managed_zeros,
               managed_empty_like=managed_empty_like,
               mem_attach_flags=types.SimpleNamespace(GLOBAL=1))
    pc = stub('pycuda', driver=drv)
    pc.autoinit = stub('pycuda.autoinit', context=_Ctx())
    pc.compiler = stub('pycuda.compiler', SourceModule=_Anything)
    pc.gpuarray = stub('pycuda.gpuarray',
                       empty=lambda n, dtype: np.empty(n, dtype))

    if root not in sys.path:
        sys.path.insert(0, root)
    import scipy.sparse.linalg  # noqa: F401  (test_kernel.py uses sp.linalg.cg)
