"""Node reordering that shrinks the octile footprint of a graph.

``rcm`` mirrors the reference's ``graphdot.graph.reorder.rcm`` (reference
graph/reorder/rcm.py:7-22): it returns a permutation for ``Graph.permute`` and
does not modify the graph.  The reference's partition-based reordering
(graph/reorder/pbr, a hypergraph partitioner around kahypar) is not rebuilt:
the solver kernels of this package gather through a CSR row index, so the
number of non-empty 8 x 8 tiles only affects the size of a packed blob, not
the work per matvec.
"""
import numpy as np


def rcm(g):
    """Reverse Cuthill-McKee permutation of a graph's nodes."""
    import scipy.sparse.csgraph
    return scipy.sparse.csgraph.reverse_cuthill_mckee(
        g.adjacency_matrix.tocsr(), symmetric_mode=True)


def octile_count(g, perm=None):
    """Number of non-empty 8 x 8 adjacency tiles, optionally after relabelling
    the nodes with ``perm`` (new index of old node ``perm[k]`` is ``k``, as in
    ``Graph.permute``)."""
    i = np.asarray(g.edges['!i']).astype(np.int64)
    j = np.asarray(g.edges['!j']).astype(np.int64)
    if perm is not None:
        inverse = np.argsort(perm)
        i, j = inverse[i], inverse[j]
    i, j = np.concatenate([i, j]), np.concatenate([j, i])
    return len(np.unique((i >> 3) * (1 << 32) + (j >> 3)))
