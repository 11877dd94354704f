"""Back-end plugin boundary of the marginalized graph kernel (reference
graphdot/kernel/marginalized/_backend.py:6-9 and _backend_factory.py:6-18).

A back end fills ``gramian`` / ``gradient`` in place for the pair ``jobs`` of
``graphs``::

    backend(graphs, node_kernel, edge_kernel, p, q, eps, ftol, gtol, jobs,
            starts, gramian, gradient, nX, nY, nJ, traits, timer)

and offers the static allocators ``array/zeros/empty`` for buffers that it can
read and write (reference _backend_cuda.py:37-47, :247-248)."""
from abc import ABC, abstractmethod


class Backend(ABC):
    @abstractmethod
    def __call__(self, graphs, node_kernel, edge_kernel, p, q, eps, ftol,
                 gtol, jobs, starts, gramian, gradient, nX, nY, nJ, traits,
                 timer):
        pass


def backend_factory(backend, *args, **kwargs):
    """``Backend`` instance, ``'b200'``/``'cuda'`` or ``'auto'``.  There is
    exactly one engine (sm_100a CUDA); there is no CPU fallback."""
    if isinstance(backend, Backend):
        return backend
    if backend in ('auto', 'cuda', 'b200'):
        from ._backend_b200 import B200Backend
        try:
            return B200Backend(*args, **kwargs)
        except Exception as e:
            if backend == 'auto':
                raise RuntimeError(f'Cannot auto-select backend: {e}')
            raise
    raise ValueError(f'Unknown backend {backend}')
