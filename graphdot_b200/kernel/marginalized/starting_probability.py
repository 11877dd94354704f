"""Starting probabilities of the random walk (reference
graphdot/kernel/marginalized/starting_probability.py:9-140).

A starting probability offers ``p(nodes) -> (values, d_values)`` on the host,
``gen_expr() -> (cxx_expr, [cxx_jacobian...])`` over the device variable ``n``
(one node), and the ``dtype``/``state`` hyper-parameter struct mirror."""
from abc import ABC, abstractmethod
from collections import namedtuple

import numpy as np


class StartingProbability(ABC):
    @abstractmethod
    def __call__(self, nodes):
        """``(p, d_p)`` for a node data frame; ``d_p`` has one row per
        hyper-parameter."""

    @abstractmethod
    def gen_expr(self):
        pass

    @property
    @abstractmethod
    def theta(self):
        pass

    @theta.setter
    @abstractmethod
    def theta(self, values):
        pass

    @property
    @abstractmethod
    def bounds(self):
        pass


class Uniform(StartingProbability):
    """Same starting probability ``p`` on every node."""

    def __init__(self, p, p_bounds=(1e-3, 1e3)):
        if not (p_bounds == 'fixed' if isinstance(p_bounds, str)
                else (isinstance(p_bounds, tuple) and len(p_bounds) == 2)):
            raise ValueError(f'invalid p_bounds {p_bounds!r}')
        self.p = p
        self.p_bounds = p_bounds

    def __call__(self, nodes):
        n = len(nodes)
        return self.p * np.ones(n), np.ones((1, n))

    def gen_expr(self):
        return 'p', ['1.f']

    dtype = np.dtype([('p', np.float32)], align=True)

    @property
    def state(self):
        return (np.float32(self.p),)

    @property
    def theta(self):
        return namedtuple('Uniform', ['p'])(self.p)

    @theta.setter
    def theta(self, values):
        self.p = values[0]

    @property
    def bounds(self):
        return (self.p_bounds,)


class Adhoc(StartingProbability):
    """``(callable over a node data frame, C++ expression over node 'n')``;
    has no trainable hyper-parameter."""

    def __init__(self, f, expr):
        self.f = f
        self.expr = expr

    def __call__(self, nodes):
        return self.f(nodes), np.empty((0, 0))

    def gen_expr(self):
        return f'({self.expr})', []

    dtype = np.dtype([('null', np.int8)], align=True)

    @property
    def state(self):
        return (np.int8(0),)

    @property
    def theta(self):
        return tuple()

    @theta.setter
    def theta(self, values):
        pass

    @property
    def bounds(self):
        return tuple()
