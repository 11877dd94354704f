"""Starting probabilities of the random walk.

Interface of the reference's starting probabilities (reference
graphdot/kernel/marginalized/starting_probability.py:9-140): an object ``p``
gives ``p(nodes) -> (values, d_values)`` on the host, ``p.gen_expr() ->
(cxx_expr, [cxx_jacobian, ...])`` over the device variable ``n`` (one node),
and the hyper-parameter plumbing the kernel front end and the back end use
(``theta``, ``bounds``, ``dtype``, ``state``).

Here the plumbing is written once: a subclass only declares the names of its
trainable scalars in ``_scalars`` (stored as attributes ``<name>`` and
``<name>_bounds``) and provides the host function and the device expression.
"""
from collections import namedtuple

import numpy as np


def _check_bounds(name, bounds):
    ok = bounds == 'fixed' if isinstance(bounds, str) else (
        isinstance(bounds, tuple) and len(bounds) == 2)
    if not ok:
        raise ValueError(f'invalid {name} {bounds!r}')
    return bounds


class StartingProbability:
    """Base class: hyper-parameter plumbing from the ``_scalars`` declaration."""

    _scalars = ()     # trainable float32 scalars, in device struct order

    def __call__(self, nodes):
        """``(p, d_p)`` for a node data frame; ``d_p`` has one row per
        hyper-parameter."""
        raise NotImplementedError

    def gen_expr(self):
        """``(C++ value expression, [C++ Jacobian expressions])``."""
        raise NotImplementedError

    # device mirror: one float32 per scalar (an int8 placeholder if none)
    @property
    def dtype(self):
        fields = [(s, np.float32) for s in self._scalars]
        return np.dtype(fields or [('null', np.int8)], align=True)

    @property
    def state(self):
        if not self._scalars:
            return (np.int8(0),)
        return tuple(np.float32(getattr(self, s)) for s in self._scalars)

    # host side: named tuple of current values / tuple of bounds
    @property
    def theta(self):
        if not self._scalars:
            return ()
        record = namedtuple(type(self).__name__, self._scalars)
        return record(*(getattr(self, s) for s in self._scalars))

    @theta.setter
    def theta(self, values):
        for s, v in zip(self._scalars, values):
            setattr(self, s, v)

    @property
    def bounds(self):
        return tuple(getattr(self, s + '_bounds') for s in self._scalars)


class Uniform(StartingProbability):
    """The same starting probability ``p`` on every node."""

    _scalars = ('p',)

    def __init__(self, p, p_bounds=(1e-3, 1e3)):
        self.p = p
        self.p_bounds = _check_bounds('p_bounds', p_bounds)

    def __call__(self, nodes):
        count = len(nodes)
        return np.full(count, self.p, dtype=float), np.ones((1, count))

    def gen_expr(self):
        return 'p', ['1.f']


class Adhoc(StartingProbability):
    """A user-supplied pair (host callable over a node data frame, C++
    expression over the node ``n``); nothing to train."""

    def __init__(self, f, expr):
        self.f, self.expr = f, expr

    def __call__(self, nodes):
        return self.f(nodes), np.empty((0, 0))

    def gen_expr(self):
        return '(' + self.expr + ')', []
