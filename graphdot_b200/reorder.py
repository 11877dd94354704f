"""Node reordering that shrinks the tile footprint of a graph.

Both functions return a permutation for ``Graph.permute`` (new index of old
node ``perm[k]`` is ``k``) and do not modify the graph, like the reference's
``graphdot.graph.reorder.rcm`` / ``pbr`` (reference graph/reorder/rcm.py:7-22,
graph/reorder/pbr/__init__.py:11-33).  They run in the native library
(``gdb_graph_reorder``, csrc/gdb_pack.cpp):

``rcm``   reverse Cuthill-McKee (the reference calls scipy).
``pbr``   the role of the reference's partition-based reordering, which
          minimises the number of non-empty 8 x 8 adjacency tiles with a
          hypergraph partitioner (kahypar, absent here): a greedy growth of
          8-node blocks -- each block is filled with the unassigned node that
          has the most neighbours inside it, a new block is seeded next to the
          previous one.

Why it matters here: the small-pair and the general kernel gather through CSR
rows and do not care about the node order; the large-pair kernel stages, per
tile row, the rows of the search direction that its elements touch, so its
shared-memory footprint and throughput follow the locality of the order
(DESIGN.md section 4.3)."""
import ctypes as C

import numpy as np

from . import native

RCM, TILES = 0, 1


def _edges(g):
    ei = np.ascontiguousarray(g.edges['!i'], dtype=np.uint32)
    ej = np.ascontiguousarray(g.edges['!j'], dtype=np.uint32)
    return ei, ej


def _reorder(g, method):
    ei, ej = _edges(g)
    n = len(g.nodes)
    perm = np.empty(n, dtype=np.uint32)
    native.check(native.load().gdb_graph_reorder(
        n, len(ei), ei.ctypes.data, ej.ctypes.data, method,
        perm.ctypes.data))
    return perm.astype(np.int64)


def rcm(g):
    """Reverse Cuthill-McKee permutation of a graph's nodes."""
    return _reorder(g, RCM)


def pbr(g):
    """Tile-minimising permutation (greedy growth of 8-node blocks)."""
    return _reorder(g, TILES)


def octile_count(g, perm=None):
    """Number of non-empty 8 x 8 adjacency tiles, optionally after relabelling
    the nodes with ``perm`` (as in ``Graph.permute``)."""
    ei, ej = _edges(g)
    p = (np.ascontiguousarray(perm, dtype=np.uint32)
         if perm is not None else None)
    out = C.c_uint64()
    native.check(native.load().gdb_graph_count_tiles(
        len(g.nodes), len(ei), ei.ctypes.data, ej.ctypes.data,
        p.ctypes.data if p is not None else None, C.byref(out)))
    return int(out.value)
