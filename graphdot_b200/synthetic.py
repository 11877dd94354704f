"""Seeded synthetic inputs of the BASELINE.json configurations (SURVEY.md
section 8(d)).  There is no network for datasets; these generators define the
workloads of bench.py and of the full-size parity tests.

C1  100 unlabeled, unweighted connected random graphs, 10-20 nodes.
C2  "molecules" of 16-24 nodes: bonded random tree + ring closures, max
    degree 4; node attributes element:int8, x:float32; edge attributes
    length:float32 and weights (the schema of reference
    graphdot/graph/_from_ase.py:34-65).
C4  Newman-Watts-Strogatz rings (k=4, p=0.05; generator family of reference
    benchmark/kernel/marginalized/time_kernel.py:20) of 200-500 nodes with a
    length-8 float32 vector feature per node.
C5  20 000 C2-style molecules.
"""
import numpy as np

from .graph import DataFrame, Graph
from .microkernel import (Constant, Convolution, KroneckerDelta,
                          SquareExponential, TensorProduct)

SEEDS = {'C1': 1001, 'C2': 2002, 'C3': 2002, 'C4': 4004, 'C5': 5005}
DEFAULT_COUNT = {'C1': 100, 'C2': 2000, 'C3': 2000, 'C4': 500, 'C5': 20000}


def _frame(cols):
    df = DataFrame()
    for k, v in cols.items():
        df[k] = v
    return df


def random_connected_graph(rng, n, p=0.3):
    perm = rng.permutation(n)
    edges = {(min(a, b), max(a, b)) for a, b in zip(perm[:-1], perm[1:])}
    iu, ju = np.triu_indices(n, 1)
    pick = rng.random(len(iu)) < p
    edges |= set(zip(iu[pick].tolist(), ju[pick].tolist()))
    e = np.array(sorted(edges), dtype=np.uint32)
    return Graph(_frame({'!i': np.arange(n, dtype=np.uint32)}),
                 _frame({'!i': e[:, 0], '!j': e[:, 1]}), title=f'er{n}')


def random_molecule(rng, n):
    deg = np.zeros(n, dtype=int)
    edges = set()
    for v in range(1, n):
        cand = [u for u in range(max(0, v - 4), v) if deg[u] < 4]
        if not cand:
            cand = [u for u in range(v) if deg[u] < 4]
        u = int(rng.choice(cand))
        edges.add((u, v))
        deg[u] += 1
        deg[v] += 1
    for _ in range(n // 6):
        for _try in range(8):
            u, v = sorted(rng.choice(n, 2, replace=False).tolist())
            if (u, v) not in edges and deg[u] < 4 and deg[v] < 4:
                edges.add((u, v))
                deg[u] += 1
                deg[v] += 1
                break
    e = np.array(sorted(edges), dtype=np.uint32)
    m = len(e)
    nodes = _frame({
        '!i': np.arange(n, dtype=np.uint32),
        'element': rng.choice(np.array([1, 6, 7, 8], dtype=np.int8), n,
                              p=[.5, .3, .1, .1]),
        'x': rng.uniform(0, 1, n).astype(np.float32),
    })
    edf = _frame({
        '!i': e[:, 0], '!j': e[:, 1],
        '!w': rng.uniform(0.5, 1.0, m).astype(np.float32),
        'length': rng.uniform(1.0, 1.6, m).astype(np.float32),
    })
    return Graph(nodes, edf, title=f'mol{n}')


def random_labeled_graph(rng, n, p_edge):
    """Connected G(n, p) graph with the molecular attribute set (test helper:
    arbitrary degrees exercise the helper-lane and overflow paths of the
    small-pair kernel)."""
    edges = {(int(rng.integers(v)), v) for v in range(1, n)}
    for u in range(n):
        for v in range(u + 1, n):
            if rng.random() < p_edge:
                edges.add((u, v))
    e = np.array(sorted(edges), dtype=np.uint32)
    m = len(e)
    nodes = _frame({
        '!i': np.arange(n, dtype=np.uint32),
        'element': rng.choice(np.array([1, 6, 7, 8], dtype=np.int8), n),
        'x': rng.uniform(0, 1, n).astype(np.float32),
    })
    edf = _frame({
        '!i': e[:, 0], '!j': e[:, 1],
        '!w': rng.uniform(0.5, 1.0, m).astype(np.float32),
        'length': rng.uniform(1.0, 1.6, m).astype(np.float32),
    })
    return Graph(nodes, edf, title=f'gnp{n}')


def newman_watts_strogatz(rng, n, k=4, p=0.05, n_feat=8):
    edges = set()
    for j in range(1, k // 2 + 1):
        for u in range(n):
            v = (u + j) % n
            edges.add((min(u, v), max(u, v)))
    for (u, v) in sorted(edges):
        if rng.random() < p:
            w = int(rng.integers(n))
            if w != u and (min(u, w), max(u, w)) not in edges:
                edges.add((min(u, w), max(u, w)))
    e = np.array(sorted(edges), dtype=np.uint32)
    feat = np.empty(n, dtype=object)
    vals = rng.normal(size=(n, n_feat)).astype(np.float32)
    for i in range(n):
        feat[i] = vals[i]
    nodes = _frame({'!i': np.arange(n, dtype=np.uint32), 'feat': feat})
    edf = _frame({'!i': e[:, 0], '!j': e[:, 1],
                  'length': rng.uniform(1.0, 1.6, len(e)).astype(np.float32)})
    return Graph(nodes, edf, title=f'nws{n}')


def make_config_graphs(config, n_graphs=None, seed=None):
    """Graphs of BASELINE configuration 'C1' .. 'C5' (first ``n_graphs``)."""
    config = config.upper()
    count = DEFAULT_COUNT[config] if n_graphs is None else n_graphs
    rng = np.random.default_rng(SEEDS[config] if seed is None else seed)
    out = []
    for g in range(count):
        if config == 'C1':
            out.append(random_connected_graph(rng, int(rng.integers(10, 21))))
        elif config in ('C2', 'C3', 'C5'):
            out.append(random_molecule(rng, int(rng.integers(16, 25))))
        elif config == 'C4':
            n = int(rng.integers(200, 501))
            out.append(newman_watts_strogatz(
                np.random.default_rng([SEEDS['C4'], g]), n))
        else:
            raise KeyError(config)
    return out


def make_config_kernel(config, **kwargs):
    """The MarginalizedGraphKernel of a BASELINE configuration."""
    from .kernel.marginalized import MarginalizedGraphKernel
    config = config.upper()
    if config == 'C1':
        kn, ke = Constant(1.0), Constant(1.0)
    elif config in ('C2', 'C3', 'C5'):
        kn = TensorProduct(element=KroneckerDelta(0.5),
                           x=SquareExponential(1.0))
        ke = TensorProduct(length=SquareExponential(0.1))
    elif config == 'C4':
        kn = TensorProduct(feat=Convolution(SquareExponential(1.0)))
        ke = TensorProduct(length=SquareExponential(0.2))
    else:
        raise KeyError(config)
    kwargs.setdefault('q', 0.05)
    kwargs.setdefault('p', 1.0)
    return MarginalizedGraphKernel(kn, ke, **kwargs)
