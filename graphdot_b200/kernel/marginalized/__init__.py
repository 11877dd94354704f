from ._kernel import MarginalizedGraphKernel
from ._backend import Backend, backend_factory

__all__ = ['MarginalizedGraphKernel', 'Backend', 'backend_factory']
