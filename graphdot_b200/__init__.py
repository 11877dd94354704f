"""graphdot_b200 — B200-native marginalized graph kernel (MLGK) engine.

Drop-in for the reference's ``graphdot.kernel.marginalized`` hot path:
``MarginalizedGraphKernel(node_kernel, edge_kernel, q, ...)(X, Y,
eval_gradient, nodal, lmin)``, ``.diag()``, ``Normalization`` and the
``microkernel`` composition language, running on hand-written sm_100a CUDA
behind a C-ABI library (include/graphdot_b200.h).
"""
from .graph import Graph

__all__ = ['Graph']
__version__ = '0.1.0'
