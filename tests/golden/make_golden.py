#!/usr/bin/env python
"""Generate golden vectors from the REFERENCE's own code (build container only).

TEST INFRASTRUCTURE.  Run once in the build container:

    python tests/golden/make_golden.py

It imports /root/reference/graphdot under ``_refshim`` and executes the
reference's own dense numpy oracle ``MLGK`` and fixtures
(reference test/kernel/marginalized/test_kernel.py:20-68 and :129-170) plus the
reference microkernels' Python ``__call__`` (graphdot/microkernel/*.py), and
writes the results as JSON next to this file.  The JSON files are committed;
this script cannot run on the GPU box (no /root/reference there).

Cross-graph known answers are obtained from the reference's single-graph oracle
through the disjoint-union identity: the product graph of H = G0 (+) G1 with
itself splits into four independent blocks, so the [G0 nodes, G1 nodes] block
of ``MLGK(H, nodal=True)`` is the nodal solution of the pair (G0, G1).
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()

from graphdot import Graph  # noqa: E402
from graphdot.minipandas import DataFrame  # noqa: E402
from graphdot import microkernel as mk  # noqa: E402

spec = importlib.util.spec_from_file_location(
    'ref_test_kernel',
    os.path.join(_refshim.REFERENCE_ROOT,
                 'test/kernel/marginalized/test_kernel.py'))
ref_test = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_test)
MLGK = ref_test.MLGK


def jsonable(v):
    if isinstance(v, np.ndarray):
        return [jsonable(x) for x in v.tolist()]
    if isinstance(v, (list, tuple)):
        return [jsonable(x) for x in v]
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (np.integer, int)):
        return int(v)
    if isinstance(v, (np.bool_, bool)):
        return bool(v)
    return v


def dump_graph(g):
    """Columns of the reference Graph as plain lists (+ numpy dtype strings)."""
    out = {}
    for part in ('nodes', 'edges'):
        df = getattr(g, part)
        cols = {}
        for key in df.columns:
            s = df[key]
            ct = s.concrete_type
            if isinstance(ct, np.dtype):
                cols[key] = {'dtype': ct.str, 'data': jsonable(np.asarray(s))}
            else:
                inner = np.asarray(s[0]).dtype.str
                cols[key] = {'dtype': 'seq:' + inner,
                             'data': [jsonable(np.asarray(x)) for x in s]}
        out[part] = cols
    out['title'] = g.title
    return out


def disjoint_union(g0, g1):
    n0 = len(g0.nodes)
    nodes = DataFrame()
    for key in g0.nodes.columns:
        if key == '!i':
            nodes[key] = np.concatenate([g0.nodes[key], g1.nodes[key] + n0])
        else:
            a, b = g0.nodes[key], g1.nodes[key]
            if isinstance(a.concrete_type, np.dtype):
                nodes[key] = np.concatenate([a, b])
            else:
                nodes[key] = list(a) + list(b)
    edges = DataFrame()
    for key in g0.edges.columns:
        a, b = g0.edges[key], g1.edges[key]
        if key in ('!i', '!j'):
            edges[key] = np.concatenate([a, b + n0])
        elif isinstance(a.concrete_type, np.dtype):
            edges[key] = np.concatenate([a, b])
        else:
            edges[key] = list(a) + list(b)
    return Graph(nodes, edges, title='union')


def mlgk_entry(G, knode, kedge, q):
    g0, g1 = G
    n0 = len(g0.nodes)
    H = disjoint_union(g0, g1)
    nodal_union = MLGK(H, knode, kedge, q, q, nodal=True)
    return {
        'q': q,
        'K00': float(MLGK(g0, knode, kedge, q, q)),
        'K11': float(MLGK(g1, knode, kedge, q, q)),
        'nodal00': jsonable(MLGK(g0, knode, kedge, q, q, nodal=True)),
        'nodal11': jsonable(MLGK(g1, knode, kedge, q, q, nodal=True)),
        'nodal01': jsonable(nodal_union[:n0, n0:]),
        'K01': float(nodal_union[:n0, n0:].sum()),
    }


def random_molecule(rng, n):
    """Small C2-schema molecule: tree + ring closures."""
    edges = set()
    for v in range(1, n):
        u = int(rng.integers(max(0, v - 4), v))
        edges.add((u, v))
    for _ in range(max(1, n // 6)):
        u, v = sorted(rng.choice(n, 2, replace=False).tolist())
        if u != v:
            edges.add((u, v))
    edges = sorted(edges)
    nodes = DataFrame({
        '!i': np.arange(n, dtype=np.uint32),
        'element': rng.choice([1, 6, 7, 8], n, p=[.5, .3, .1, .1]
                              ).astype(np.int8),
        'x': rng.uniform(0, 1, n).astype(np.float32),
    })
    i, j = np.array(edges, dtype=np.uint32).T
    e = DataFrame({
        '!i': i, '!j': j,
        '!w': rng.uniform(0.5, 1.0, len(i)).astype(np.float32),
        'length': rng.uniform(1.0, 1.6, len(i)).astype(np.float32),
    })
    return Graph(nodes, e, title=f'mol{n}')


def main():
    out = {'_provenance': 'reference MLGK() test/kernel/marginalized/'
                          'test_kernel.py:20-68 run under tests/golden/'
                          '_refshim.py; graphdot 0.8.1',
           'cases': {}}
    for name, case in ref_test.case_dict.items():
        G = case['graphs']
        knode, kedge = case['knode'], case['kedge']
        out['cases'][name] = {
            'graphs': [dump_graph(g) for g in G],
            'knode': repr(knode), 'kedge': repr(kedge),
            'entries': [mlgk_entry(G, knode, kedge, q) for q in case['q']],
        }

    # random molecular pairs with the C2 kernels.  MLGK() calls
    # scipy CG with its default rtol=1e-5, which leaves ~1e-4 relative error on
    # these larger systems; tighten the solver (not the reference's assembly)
    # so the vectors pin the oracle to 1e-9.
    import functools
    import scipy.sparse.linalg as spla
    loose_cg = spla.cg
    spla.cg = functools.partial(loose_cg, rtol=1e-14, maxiter=100000)
    out['_provenance'] += ('; cases molecular* solved with scipy cg '
                           'rtol=1e-14 instead of the default 1e-5')
    rng = np.random.default_rng(2002)
    knode = mk.TensorProduct(element=mk.KroneckerDelta(0.5),
                             x=mk.SquareExponential(1.0))
    kedge = mk.TensorProduct(length=mk.SquareExponential(0.1))
    G = [random_molecule(rng, 7), random_molecule(rng, 10)]
    out['cases']['molecular'] = {
        'graphs': [dump_graph(g) for g in G],
        'knode': repr(knode), 'kedge': repr(kedge),
        'entries': [mlgk_entry(G, knode, kedge, q) for q in (0.05, 0.2)],
    }
    G = [random_molecule(rng, 17), random_molecule(rng, 12)]
    out['cases']['molecular-multitile'] = {
        'graphs': [dump_graph(g) for g in G],
        'knode': repr(knode), 'kedge': repr(kedge),
        'entries': [mlgk_entry(G, knode, kedge, q) for q in (0.05,)],
    }
    spla.cg = loose_cg
    with open(os.path.join(HERE, 'mlgk_reference.json'), 'w') as f:
        json.dump(out, f, indent=1)

    # ---- microkernel values / Jacobians from the reference's __call__ ----
    rng = np.random.default_rng(7)
    mkout = {'_provenance': 'reference graphdot/microkernel __call__(x, y, '
                            'jac=True), graphdot 0.8.1', 'items': []}

    def record(expr, kernel, pairs):
        vals = []
        for x, y in pairs:
            f, j = kernel(x, y, jac=True)
            vals.append({'x': jsonable(x), 'y': jsonable(y),
                         'f': float(f), 'jac': jsonable(np.asarray(j, float))})
        mkout['items'].append({'expr': expr, 'repr': repr(kernel),
                               'theta': jsonable(list(ref_flatten(kernel.theta))),
                               'minmax': jsonable(list(kernel.minmax)),
                               'samples': vals})

    from graphdot.util.iterable import flatten as ref_flatten
    scalars = [(float(a), float(b)) for a, b in rng.normal(size=(6, 2))]
    ints = [(1, 1), (1, 2), (6, 8), (0, 0)]
    record('Constant(0.7)', mk.Constant(0.7), scalars[:2])
    record('KroneckerDelta(0.3)', mk.KroneckerDelta(0.3), ints)
    record('SquareExponential(0.8)', mk.SquareExponential(0.8), scalars)
    record('RationalQuadratic(0.9, 1.7)',
           mk.RationalQuadratic(0.9, 1.7), scalars)
    record('SquareExponential(1.0) + 0.01',
           mk.SquareExponential(1.0) + 0.01, scalars)
    record('KroneckerDelta(0.5) * SquareExponential(2.0)',
           mk.KroneckerDelta(0.5) * mk.SquareExponential(2.0),
           [(1.0, 1.0), (1.0, 2.5), (0.3, -0.4)])
    record('(SquareExponential(1.0) + 0.5) ** 2.0',
           (mk.SquareExponential(1.0) + 0.5) ** 2.0, scalars[:3])
    record('(KroneckerDelta(0.4) * 0.5 + 0.25).normalized',
           (mk.KroneckerDelta(0.4) * 0.5 + 0.25).normalized, ints)
    dicts = [({'a': 1, 'b': 0.5}, {'a': 1, 'b': 1.5}),
             ({'a': 2, 'b': -0.5}, {'a': 1, 'b': 0.25}),
             ({'a': 3, 'b': 0.0}, {'a': 3, 'b': 0.0})]
    record("TensorProduct(a=KroneckerDelta(0.3), b=SquareExponential(1.0))",
           mk.TensorProduct(a=mk.KroneckerDelta(0.3),
                            b=mk.SquareExponential(1.0)), dicts)
    record("Additive(a=KroneckerDelta(0.3), b=SquareExponential(0.05))"
           ".normalized",
           mk.Additive(a=mk.KroneckerDelta(0.3),
                       b=mk.SquareExponential(0.05)).normalized, dicts)
    record("TensorProduct(a=KroneckerDelta(0.3), "
           "b=SquareExponential(1.) + 0.01).normalized",
           mk.TensorProduct(a=mk.KroneckerDelta(0.3),
                            b=mk.SquareExponential(1.) + 0.01).normalized,
           dicts)
    seqs = [((5, 6), (3,)), ((2, 3, 4), (3, 4)), ((3,), (3,))]
    record('Convolution(KroneckerDelta(0.3))',
           mk.Convolution(mk.KroneckerDelta(0.3)), seqs)
    record('Convolution(SquareExponential(1.0), mean=False)',
           mk.Convolution(mk.SquareExponential(1.0), mean=False), seqs)
    vecs = [((1.0, 2.0, 3.0), (0.5, -1.0, 2.0)), ((0.0, 1.0), (1.0, 0.0))]
    record('DotProduct()', mk.DotProduct(), vecs)
    record('Product()', mk.Product(), scalars[:3])
    with open(os.path.join(HERE, 'microkernel_reference.json'), 'w') as f:
        json.dump(mkout, f, indent=1)
    print('wrote', len(out['cases']), 'MLGK cases and',
          len(mkout['items']), 'microkernel items')


if __name__ == '__main__':
    main()
