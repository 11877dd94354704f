"""Graph container accepted by the MLGK engine.

Host-side mirror of the reference's input type so that the same graphs can be
fed to either implementation (reference graphdot/graph/__init__.py:40-249,
graphdot/minipandas/dataframe.py, graphdot/minipandas/series.py,
graphdot/graph/_from_networkx.py).  Only what the marginalized-kernel path
needs is provided: column store, ``rowtype``, ``from_networkx``,
``unify_datatype``, ``has_unified_types``, ``permute`` and the volatile
``cookie`` used by back ends to cache device-side data.
"""
import copy as _copy
import itertools as _it
from collections import namedtuple

import numpy as np

__all__ = ['Graph', 'DataFrame', 'Series']


# --------------------------------------------------------------------------
# element-type inference (reference graphdot/codegen/typetool.py:27-123)
# --------------------------------------------------------------------------
def _is_scalar_dtype(t):
    return isinstance(t, np.dtype) and t.kind != 'O' and t.names is None


def _signed(t):
    if isinstance(t, np.dtype) and t.kind == 'u':
        return np.promote_types(t, np.int8)
    return t


def _merge(types, coerce=True):
    """Smallest type able to hold every type in ``types``.  Non-numpy python
    classes (list, tuple, ndarray, str ...) only merge with themselves."""
    acc = None
    for t in types:
        t = _signed(t)
        if acc is None:
            acc = t
        elif acc != t:
            if not coerce:
                return None
            acc = np.promote_types(acc, t)
    if isinstance(acc, np.dtype) and acc.kind == 'f':
        acc = np.promote_types(acc, np.float32)
    return acc


def smallest_type_of_values(values, coerce=True):
    return _merge((np.min_scalar_type(v) if np.isscalar(v) else type(v)
                   for v in values), coerce)


def smallest_type_of_types(types, coerce=True):
    return _merge(types, coerce)


class Series(np.ndarray):
    """1-D column.  ``concrete_type`` is a numpy dtype for scalar columns and
    the python class of the elements (list/tuple/ndarray) otherwise."""

    def __new__(cls, values):
        if isinstance(values, np.ndarray):
            obj = values.view(cls)
            if _is_scalar_dtype(obj.dtype):
                obj._concrete_type = obj.dtype
            else:
                kinds = {type(v) for v in values}
                obj._concrete_type = kinds.pop() if len(kinds) == 1 else None
            return obj
        values = list(values)
        t = smallest_type_of_values(values)
        dtype = t if _is_scalar_dtype(t) else np.dtype(object)
        obj = np.empty(len(values), dtype=dtype).view(cls)
        obj[:] = values
        obj._concrete_type = t
        return obj

    def __array_finalize__(self, parent):
        if parent is not None and not hasattr(self, '_concrete_type'):
            ct = getattr(parent, '_concrete_type', None)
            self._concrete_type = self.dtype if _is_scalar_dtype(self.dtype) \
                else ct

    @property
    def concrete_type(self):
        return self._concrete_type

    def __repr__(self):
        return np.array2string(np.asarray(self), separator=',',
                               max_line_width=10**9)

    def __reduce__(self):
        fn, args, state = super().__reduce__()
        return fn, args, (state, self._concrete_type)

    def __setstate__(self, state):
        base, ct = state
        super().__setstate__(base)
        self._concrete_type = ct


class DataFrame:
    """Ordered mapping ``column name -> Series`` of equal-length columns."""

    def __init__(self, data=None):
        self._data = {}
        if data is not None:
            for key, value in dict(data).items():
                self[key] = value

    # -- mapping protocol --------------------------------------------------
    def __setitem__(self, key, value):
        self._data[key] = value if isinstance(value, Series) else Series(value)
        self._sig = None    # cached type signature (see Graph.has_unified_types)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._data[key]
        sel = np.asarray(key)
        if sel.dtype == np.bool_:
            return DataFrame({k: v[sel] for k, v in self._data.items()})
        return DataFrame({k: self._data[k] for k in key})

    def __getattr__(self, name):
        data = self.__dict__.get('_data')
        if data is not None and name in data:
            return data[name]
        raise AttributeError(f'DataFrame has no column {name!r}')

    def __contains__(self, key):
        return key in self._data

    def __iter__(self):
        return iter(self._data)

    def __len__(self):
        return max((len(c) for c in self._data.values()), default=0)

    def __repr__(self):
        return repr(self._data)

    @property
    def columns(self):
        return list(self._data)

    # -- row views ---------------------------------------------------------
    def rowtype(self, pack=True):
        """Aligned struct dtype of one row; with ``pack`` the fields are
        ordered by decreasing item size (stable), which is the field order of
        the device-side ``node_t``/``edge_t`` (reference
        graphdot/minipandas/dataframe.py:55-63)."""
        names = self.columns
        dts = [np.dtype(self[k].concrete_type).newbyteorder('=')
               for k in names]
        order = range(len(names))
        if pack:
            order = np.argsort([-d.itemsize for d in dts], kind='stable')
        return np.dtype([(names[i], dts[i]) for i in order], align=True)

    def rows(self, rowname='row'):
        visible = [k for k in self._data if k.isidentifier()]
        base = namedtuple(rowname, visible)

        class Row(base):
            __slots__ = ()

            def __getitem__(self, key):
                if isinstance(key, str):
                    return getattr(self, key)
                return base.__getitem__(self, key)

        Row.__name__ = rowname
        cols = [self._data[k] for k in visible]
        for i in range(len(self)):
            yield Row(*[c[i] for c in cols])

    def itertuples(self, tuplename='tuple'):
        yield from self.rows(tuplename)

    def iterrows(self, rowname='row'):
        yield from enumerate(self.rows(rowname))

    def copy(self, deep=False):
        if deep:
            return DataFrame({k: Series(np.copy(v))
                              for k, v in self._data.items()})
        return DataFrame(self._data)

    def drop(self, keys, inplace=False):
        if inplace:
            for k in keys:
                self._data.pop(k, None)
            self._sig = None
            return None
        return self[[k for k in self.columns if k not in keys]]

    def to_pandas(self):
        import pandas as pd
        return pd.DataFrame({k: np.asarray(v) for k, v in self._data.items()})


class VolatileCookie(dict):
    """Per-graph cache for back ends; intentionally dropped by pickling and
    deep copies (reference graphdot/util/cookie.py:5-12).  ``epoch`` counts
    the invalidations of ANY graph's cache (``Graph.permute(inplace=True)``,
    ``Graph.unify_datatype``): callers that memoise work over a whole list of
    graphs compare it instead of walking the list."""

    epoch = 0

    def clear(self):
        VolatileCookie.epoch += 1
        super().clear()

    def pop(self, *args):
        VolatileCookie.epoch += 1
        return super().pop(*args)

    def __delitem__(self, key):
        VolatileCookie.epoch += 1
        super().__delitem__(key)

    def __reduce__(self):
        return (VolatileCookie, ())

    def __deepcopy__(self, memo):
        return VolatileCookie()


class Graph:
    """A graph as two data frames.

    ``nodes`` needs the column ``!i`` (node index 0..n-1); ``edges`` needs
    ``!i`` and ``!j`` (end points) and may carry ``!w`` (weights).  Every other
    column is a node/edge attribute visible to the microkernels.
    """

    def __init__(self, nodes, edges, title=''):
        self.title = str(title)
        self.nodes = nodes if isinstance(nodes, DataFrame) else DataFrame(nodes)
        self.edges = edges if isinstance(edges, DataFrame) else DataFrame(edges)
        if '!i' not in self.nodes:
            raise ValueError("nodes need an '!i' column")
        if '!i' not in self.edges or '!j' not in self.edges:
            raise ValueError("edges need '!i' and '!j' columns")

    def __repr__(self):
        return (f'{type(self).__name__}(nodes={self.nodes!r}, '
                f'edges={self.edges!r}, title={self.title!r})')

    @property
    def cookie(self):
        c = self.__dict__.get('_cookie')
        if c is None:
            c = self.__dict__['_cookie'] = VolatileCookie()
        return c

    def copy(self, deep=False):
        g = type(self)(self.nodes.copy(deep), self.edges.copy(deep),
                       self.title)
        for k, v in self.__dict__.items():
            if k not in ('nodes', 'edges', 'title', '_cookie'):
                g.__dict__[k] = _copy.deepcopy(v) if deep else v
        return g

    def permute(self, perm, inplace=False):
        """Relabel nodes: new index of old node ``perm[k]`` is ``k``."""
        if inplace:
            g = self
            g.cookie.clear()
        else:
            g = self.copy(deep=True)
        inverse = np.argsort(perm)
        for df, cols in ((g.nodes, ('!i',)), (g.edges, ('!i', '!j'))):
            for c in cols:
                df[c][:] = inverse[df[c]]
        return g

    @property
    def adjacency_matrix(self):
        import scipy.sparse
        n = len(self.nodes)
        i, j = np.asarray(self.edges['!i']), np.asarray(self.edges['!j'])
        w = (np.asarray(self.edges['!w']) if '!w' in self.edges
             else np.ones(len(i)))
        a = scipy.sparse.coo_matrix((w, (i, j)), shape=(n, n))
        return a + a.T

    @property
    def laplacian(self):
        import scipy.sparse
        a = self.adjacency_matrix
        return scipy.sparse.diags(np.asarray(a.sum(axis=0)).ravel(), 0) - a

    # -- type handling -------------------------------------------------------
    @staticmethod
    def has_unified_types(graphs):
        """True, or ``(component, first, offender)`` for the first mismatch."""
        graphs = list(graphs)
        first = graphs[0]

        def signature(frame):
            # (column, element type) pairs: equal signatures <=> equal
            # rowtype().  Cached on the frame (reset by every column
            # assignment / drop): 2000 graphs are checked in well under a
            # millisecond instead of building 4000 numpy struct dtypes.
            sig = frame.__dict__.get('_sig')
            if sig is None:
                sig = frame.__dict__['_sig'] = tuple(
                    (k, c.concrete_type) for k, c in frame._data.items())
            return sig

        try:
            nt, et = signature(first.nodes), signature(first.edges)
            for g in graphs:
                if signature(g.nodes) != nt:
                    break
                if signature(g.edges) != et:
                    break
            else:
                return True
        except AttributeError:      # foreign frame types: the generic way
            pass
        nt, et = first.nodes.rowtype(), first.edges.rowtype()
        for g in graphs:
            if g.nodes.rowtype() != nt:
                return ('nodes', first, g)
            if g.edges.rowtype() != et:
                return ('edges', first, g)
        return True

    @classmethod
    def unify_datatype(cls, graphs, inplace=False):
        """Give each attribute one dtype across all graphs (the smallest that
        holds every value); sequence attributes get a common element type."""
        for g in graphs:
            g.cookie.clear()
        if not inplace:
            graphs = [g.copy(deep=False) for g in graphs]
        for part in ('nodes', 'edges'):
            frames = [getattr(g, part) for g in graphs]
            keys = set(frames[0].columns)
            for g, f in zip(graphs, frames):
                if set(f.columns) != keys:
                    raise TypeError(
                        f'Graph {g.title!r}: {part} attributes '
                        f'{set(f.columns)} do not match {keys}.')
            for key in frames[0].columns:
                types = [f[key].concrete_type for f in frames]
                t = smallest_type_of_types(types)
                if t == np.dtype(object) or t is object:
                    t = smallest_type_of_types(types, coerce=False)
                if t is None:
                    raise TypeError(f'Cannot unify attribute {key!r}: mixed '
                                    'object types.')
                if _is_scalar_dtype(t):
                    for f in frames:
                        f[key] = np.asarray(f[key]).astype(t)
                elif t in (list, tuple, np.ndarray):
                    inner = smallest_type_of_values(
                        _it.chain.from_iterable(
                            _it.chain.from_iterable(f[key] for f in frames)))
                    if inner is None:
                        raise TypeError('No common element type for '
                                        f'attribute {key!r}.')
                    for f in frames:
                        col = np.empty(len(f[key]), dtype=object)
                        for k, seq in enumerate(f[key]):
                            col[k] = np.array(seq, dtype=inner)
                        f[key] = col
        if not inplace:
            return graphs

    # -- converters ------------------------------------------------------------
    @classmethod
    def from_networkx(cls, graph, weight=None):
        """Convert an undirected NetworkX graph with homogeneous attributes;
        ``weight`` names the edge attribute holding edge weights."""
        import networkx as nx
        labels = list(graph.nodes)
        if (not all(isinstance(x, (int, np.integer)) for x in labels)
                or sorted(labels) != list(range(len(labels)))):
            graph = nx.convert_node_labels_to_integers(graph)
        title = graph.graph.get('title', '')

        node_keys = None
        for idx, attrs in graph.nodes.items():
            keys = sorted(attrs)
            if node_keys is None:
                node_keys = keys
            elif keys != node_keys:
                raise TypeError(f'Node {idx} attributes {keys} inconsistent '
                                f'with {node_keys}')
        # node ids in ITERATION order (the attribute columns below follow the
        # same order; the packer sorts rows by '!i').  The reference numbers
        # them 0..n-1 regardless (reference graph/_from_networkx.py:49), which
        # misplaces attributes when integer labels are not iterated sorted.
        nodes = DataFrame({'!i': np.array(list(graph.nodes), dtype=np.uint32)})
        for key in node_keys or []:
            nodes[key] = [a[key] for a in graph.nodes.values()]

        if graph.number_of_edges() == 0:
            raise RuntimeError(f'Graph {graph} has no edges.')
        edge_keys = None
        for ij, attrs in graph.edges.items():
            keys = sorted(attrs)
            if edge_keys is None:
                edge_keys = keys
            elif keys != edge_keys:
                raise TypeError(f'Edge {ij} attributes {keys} inconsistent '
                                f'with {edge_keys}')
        edges = DataFrame()
        ei, ej = zip(*graph.edges.keys())
        edges['!i'], edges['!j'] = ei, ej
        if weight is not None:
            edges['!w'] = [a[weight] for a in graph.edges.values()]
        for key in edge_keys:
            if key != weight:
                edges[key] = [a[key] for a in graph.edges.values()]
        return cls(nodes, edges, title=title)

    @classmethod
    def from_columns(cls, nodes, edges, title=''):
        """Build from ``{column: {'dtype': str, 'data': list}}`` dictionaries
        (the layout of tests/golden/*.json)."""
        def frame(cols):
            df = DataFrame()
            for key, col in cols.items():
                dt = col['dtype']
                if dt.startswith('seq:'):
                    arr = np.empty(len(col['data']), dtype=object)
                    for k, seq in enumerate(col['data']):
                        arr[k] = np.array(seq, dtype=np.dtype(dt[4:]))
                    df[key] = arr
                else:
                    df[key] = np.array(col['data'], dtype=np.dtype(dt))
            return df
        return cls(frame(nodes), frame(edges), title=title)

    def to_networkx(self):
        import networkx as nx
        g = nx.Graph(title=self.title)
        node_keys = [k for k in self.nodes.columns if k != '!i']
        for row in range(len(self.nodes)):
            g.add_node(int(self.nodes['!i'][row]),
                       **{k: self.nodes[k][row] for k in node_keys})
        edge_keys = [k for k in self.edges.columns if k not in ('!i', '!j')]
        for row in range(len(self.edges)):
            g.add_edge(int(self.edges['!i'][row]), int(self.edges['!j'][row]),
                       **{k: self.edges[k][row] for k in edge_keys})
        return g
