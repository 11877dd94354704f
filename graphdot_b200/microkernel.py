"""Microkernels: positive-definite similarity functions between node / edge
features, composable with ``+ * **``, ``TensorProduct``, ``Additive``,
``Convolution`` and ``.normalized``.

Python surface of the reference's composition language (reference
graphdot/microkernel/_base.py:16-167 for the protocol, :170-330 operator
kernels, :333-385 Constant, :388-478 Normalize, :481-730 sympy kernels;
composite.py:10-131; convolution.py:10-96; kronecker_delta.py:9-72;
dotproduct.py:8-53; product.py:8-42).  Every kernel offers

* ``k(x, y, jac=False)``      – host evaluation (used by the CPU oracle),
* ``k.gen_expr(x, y, scope)`` – a CUDA C++ value expression and one Jacobian
  expression per hyper-parameter, spliced into the NVRTC solver template
  (graphdot_b200/csrc/mlgk_solver.cuh).  The dialect is the reference's:
  hyper-parameters are members of nested structs addressed through
  ``scope``, helpers are ``graphdot::ipow<N>``, ``graphdot::ripow<N>``,
  ``normalize``, ``normalize_jacobian``, ``convolution<mean>``,
  ``convolution_jacobian<mean>`` and ``dotproduct``,
* ``k.dtype`` / ``k.state``   – aligned numpy struct layout of the
  hyper-parameters and the matching nested value tuple (what the reference
  obtains from ``cpptype``, graphdot/codegen/cpptool.py:9-101),
* ``theta`` / ``bounds`` / ``minmax`` / ``name`` / ``repr``.
"""
import math
from abc import ABC, abstractmethod
import functools
from collections import namedtuple

import numpy as np

__all__ = ['MicroKernel', 'Product', 'Constant', 'KroneckerDelta',
           'SquareExponential', 'RationalQuadratic', 'Normalize', 'Composite',
           'TensorProduct', 'Additive', 'Convolution', 'DotProduct']


def _named(typename, fields):
    return _named_cached(typename, tuple(fields))


@functools.lru_cache(maxsize=None)
def _named_cached(typename, fields):
    # one class per (name, fields): building a namedtuple costs 50 us, and
    # every `.theta` / `.hyperparameters` access asks for one
    base = namedtuple(typename, fields)

    def __repr__(self):
        lines = []
        for f, v in zip(fields, self):
            if isinstance(v, tuple):
                body = repr(v).replace('\n', '\n\t')
                lines.append(f'{f} : {type(v).__name__}\n\t{body}')
            else:
                lines.append(f'{f} : {v!r}')
        return '\n'.join(lines)

    return type(typename, (base,), {'__slots__': (), '__repr__': __repr__})


def _check_bounds(owner, hyper, bounds):
    ok = bounds == 'fixed' if isinstance(bounds, str) else (
        isinstance(bounds, tuple) and len(bounds) == 2)
    if not ok:
        raise ValueError(f'Bounds of hyperparameter {hyper} of {owner} must '
                         f'be a 2-tuple or "fixed", got {bounds!r}.')


class MicroKernel(ABC):
    """Abstract base of all microkernels."""

    # -- protocol ----------------------------------------------------------
    @property
    @abstractmethod
    def name(self):
        """Name of the kernel."""

    @abstractmethod
    def __call__(self, x, y, jac=False):
        """Value (and Jacobian w.r.t. hyper-parameters if ``jac``)."""

    @abstractmethod
    def __repr__(self):
        pass

    @abstractmethod
    def gen_expr(self, x, y, theta_scope=''):
        """``(value_expr, [jacobian_expr, ...])`` as CUDA C++ strings over
        the variables named ``x`` and ``y``; hyper-parameters are addressed
        as ``theta_scope + name``."""

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

    @property
    @abstractmethod
    def minmax(self):
        pass

    @abstractmethod
    def _layout(self):
        """List of ``(member, np.float32-like dtype | MicroKernel)``."""

    # -- hyper-parameter struct mirror --------------------------------------
    @property
    def dtype(self):
        fields = []
        for key, what in self._layout():
            fields.append((key, what.dtype if isinstance(what, MicroKernel)
                           else np.dtype(what)))
        return np.dtype(fields, align=True)

    @property
    def state(self):
        out = []
        for key, what in self._layout():
            if isinstance(what, MicroKernel):
                out.append(what.state)
            else:
                out.append(np.dtype(what).type(getattr(self, key)))
        return tuple(out)

    # -- composition ----------------------------------------------------------
    @property
    def normalized(self):
        return Normalize(self)

    def __add__(self, other):
        return _Add(self, other)

    def __radd__(self, other):
        return _Add(other, self)

    def __mul__(self, other):
        return _Multiply(self, other)

    def __rmul__(self, other):
        return _Multiply(other, self)

    def __pow__(self, exponent):
        return _Exponentiation(self, exponent)

    @staticmethod
    def from_sympy(name, desc, expr, vars, *hyperparameter_specs,
                   minmax=(0, 1)):
        """Create a scalar microkernel class from a SymPy expression.  Specs
        are ``symbol`` | ``(symbol, dtype[, lb, ub][, doc])``."""
        return _from_sympy(name, desc, expr, vars, *hyperparameter_specs,
                           minmax=minmax)


def _as_kernel(k):
    return Constant(k) if np.isscalar(k) else k


# ----------------------------------------------------------------------------
# binary operator kernels
# ----------------------------------------------------------------------------
class _BinaryExpr(MicroKernel):
    opstr = '?'
    _name = '?'

    def __init__(self, k1, k2):
        self.k1 = _as_kernel(k1)
        self.k2 = _as_kernel(k2)

    @property
    def name(self):
        return self._name

    def __repr__(self):
        return f'{self.k1!r} {self.opstr} {self.k2!r}'

    def _layout(self):
        return [('k1', self.k1), ('k2', self.k2)]

    @property
    def theta(self):
        return _named(self.name, ['lhs', 'rhs'])(self.k1.theta, self.k2.theta)

    @theta.setter
    def theta(self, values):
        self.k1.theta, self.k2.theta = values[0], values[1]

    @property
    def bounds(self):
        return (self.k1.bounds, self.k2.bounds)

    def _sub_exprs(self, x, y, scope):
        f1, j1 = self.k1.gen_expr(x, y, scope + 'k1.')
        f2, j2 = self.k2.gen_expr(x, y, scope + 'k2.')
        return f1, j1, f2, j2

    def _sub_values(self, x, y):
        f1, j1 = self.k1(x, y, True)
        f2, j2 = self.k2(x, y, True)
        return f1, np.asarray(j1, float), f2, np.asarray(j2, float)


class _Add(_BinaryExpr):
    opstr, _name = '+', 'Add'

    def __call__(self, x, y, jac=False):
        if not jac:
            return self.k1(x, y) + self.k2(x, y)
        f1, j1, f2, j2 = self._sub_values(x, y)
        return f1 + f2, np.concatenate([j1, j2])

    def gen_expr(self, x, y, theta_scope=''):
        f1, j1, f2, j2 = self._sub_exprs(x, y, theta_scope)
        return f'({f1} + {f2})', j1 + j2

    @property
    def minmax(self):
        (a, b), (c, d) = self.k1.minmax, self.k2.minmax
        return (a + c, b + d)


class _Multiply(_BinaryExpr):
    opstr, _name = '*', 'Multiply'

    def __call__(self, x, y, jac=False):
        if not jac:
            return self.k1(x, y) * self.k2(x, y)
        f1, j1, f2, j2 = self._sub_values(x, y)
        return f1 * f2, np.concatenate([j1 * f2, f1 * j2])

    def gen_expr(self, x, y, theta_scope=''):
        f1, j1, f2, j2 = self._sub_exprs(x, y, theta_scope)
        return (f'({f1} * {f2})',
                [f'({j} * {f2})' for j in j1] + [f'({f1} * {j})' for j in j2])

    @property
    def minmax(self):
        (a, b), (c, d) = self.k1.minmax, self.k2.minmax
        return (a * c, b * d)


class _Exponentiation(_BinaryExpr):
    opstr, _name = '**', 'Exponentiation'

    def __init__(self, k1, exponent):
        if np.isscalar(exponent):
            exponent = Constant(exponent)
        elif not (isinstance(exponent, MicroKernel)
                  and exponent.name == 'Constant'):
            raise ValueError('Exponent must be a constant or a Constant '
                             f'microkernel, got {exponent!r}.')
        super().__init__(k1, exponent)

    def __call__(self, x, y, jac=False):
        if not jac:
            return self.k1(x, y) ** self.k2(x, y)
        f1, j1, f2, j2 = self._sub_values(x, y)
        return f1 ** f2, np.concatenate([f2 * f1 ** (f2 - 1) * j1,
                                         f1 ** f2 * np.log(f1) * j2])

    def gen_expr(self, x, y, theta_scope=''):
        f1, j1, f2, j2 = self._sub_exprs(x, y, theta_scope)
        return (f'__powf({f1}, {f2})',
                [f'({f2} * __powf({f1}, {f2} - 1) * {j})' for j in j1] +
                [f'(__powf({f1}, {f2}) * __logf({f1}) * {j})' for j in j2])

    @property
    def minmax(self):
        (a, b), (c, d) = self.k1.minmax, self.k2.minmax
        return (a ** c, b ** d)


# ----------------------------------------------------------------------------
# elementary kernels
# ----------------------------------------------------------------------------
class _ConstantKernel(MicroKernel):
    def __init__(self, c, c_bounds='fixed'):
        _check_bounds('Constant', 'c', c_bounds)
        self.c = float(c)
        self.c_bounds = c_bounds

    @property
    def name(self):
        return 'Constant'

    def __call__(self, x, y, jac=False):
        return (self.c, np.ones(1)) if jac else self.c

    def __repr__(self):
        return f'Constant({self.c})'

    def gen_expr(self, x, y, theta_scope=''):
        return f'{theta_scope}c', ['1.0f']

    def _layout(self):
        return [('c', np.float32)]

    @property
    def theta(self):
        return _named('Constant', ['c'])(self.c)

    @theta.setter
    def theta(self, values):
        self.c = values[0]

    @property
    def bounds(self):
        return (self.c_bounds,)

    @property
    def minmax(self):
        return (self.c, self.c)


def Constant(c, c_bounds='fixed'):
    r"""No-op microkernel :math:`k(\cdot,\cdot) \equiv c`, typically used as
    an adjustable weight."""
    return _ConstantKernel(c, c_bounds)


class _KroneckerDeltaKernel(MicroKernel):
    def __init__(self, h, h_bounds=(1e-3, 1)):
        _check_bounds('KroneckerDelta', 'h', h_bounds)
        self.h = float(h)
        self.h_bounds = h_bounds

    @property
    def name(self):
        return 'KroneckerDelta'

    def __call__(self, x, y, jac=False):
        same = bool(x == y)
        f = 1.0 if same else self.h
        return (f, np.array([0.0 if same else 1.0])) if jac else f

    def __repr__(self):
        return f'KroneckerDelta({self.h})'

    def gen_expr(self, x, y, theta_scope=''):
        return (f'({x} == {y} ? 1.0f : {theta_scope}h)',
                [f'({x} == {y} ? 0.0f : 1.0f)'])

    def _layout(self):
        return [('h', np.float32)]

    @property
    def theta(self):
        return _named('KroneckerDelta', ['h'])(self.h)

    @theta.setter
    def theta(self, values):
        self.h = values[0]

    @property
    def bounds(self):
        return (self.h_bounds,)

    @property
    def minmax(self):
        return (self.h, 1)


def KroneckerDelta(h, h_bounds=(1e-3, 1)):
    r"""1 if the two features compare equal, ``h`` in (0, 1) otherwise."""
    return _KroneckerDeltaKernel(h, h_bounds)


class _ProductKernel(MicroKernel):
    """Direct product between scalar features (the edge-weight pseudo
    kernel)."""

    @property
    def name(self):
        return 'Product'

    def __call__(self, x, y, jac=False):
        return (x * y, np.array([])) if jac else x * y

    def __repr__(self):
        return 'Product()'

    def gen_expr(self, x, y, theta_scope=''):
        return f'({x} * {y})', []

    def _layout(self):
        return []

    @property
    def theta(self):
        return tuple()

    @theta.setter
    def theta(self, values):
        pass

    @property
    def bounds(self):
        return tuple()

    @property
    def minmax(self):
        return (None, None)


def Product():
    return _ProductKernel()


class _DotProductKernel(MicroKernel):
    @property
    def name(self):
        return 'DotProduct'

    def __call__(self, x, y, jac=False):
        f = np.asarray(x) @ np.asarray(y)
        return (f, []) if jac else f

    def __repr__(self):
        return 'DotProduct()'

    def gen_expr(self, x, y, theta_scope=''):
        return f'dotproduct({x}, {y})', []

    def _layout(self):
        return []

    @property
    def theta(self):
        return tuple()

    @theta.setter
    def theta(self, values):
        pass

    @property
    def bounds(self):
        return tuple()

    @property
    def minmax(self):
        return (0, np.inf)


def DotProduct():
    """Inner product between two equal-length vector features."""
    return _DotProductKernel()


class _ScalarKernel(MicroKernel):
    """Kernel on scalar features with named float hyper-parameters.
    Subclasses set ``_kname``, ``_hypers`` (ordered ``{name: (dtype,
    default_bounds|None)}``), ``_minmax`` and implement ``_f`` (value and
    Jacobian in python) and ``_cxx`` (value and Jacobian strings)."""
    _kname = ''
    _hypers = {}
    _minmax = (0, 1)

    def __init__(self, *args, **kwargs):
        self._values, self._bounds = {}, {}
        names = list(self._hypers)
        if len(args) > len(names):
            raise TypeError(f'{self._kname} takes at most {len(names)} '
                            'positional hyperparameters')
        for key, v in zip(names, args):
            self._values[key] = v
        for key in names:
            if key in kwargs:
                self._values[key] = kwargs[key]
            if key not in self._values:
                raise KeyError(f'Hyperparameter {key} not provided for '
                               f'{self._kname}')
            b = kwargs.get(f'{key}_bounds', self._hypers[key][1])
            if b is None:
                raise KeyError(f'Bounds for hyperparameter {key} of '
                               f'microkernel {self._kname} not set, and no '
                               'defaults were given.')
            _check_bounds(self._kname, key, b)
            self._bounds[key] = b

    def __getattr__(self, key):
        vals = self.__dict__.get('_values')
        if vals is not None and key in vals:
            return vals[key]
        raise AttributeError(key)

    @property
    def name(self):
        return self._kname

    def __call__(self, x, y, jac=False):
        f, j = self._f(x, y, *self._values.values())
        return (f, np.asarray(j, float)) if jac else f

    def __repr__(self):
        vals = ', '.join(f'{k}={v}' for k, v in self._values.items())
        bnds = ', '.join(f'{k}_bounds={v}' for k, v in self._bounds.items())
        return f'{self._kname}({vals}, {bnds})'

    def gen_expr(self, x, y, theta_scope=''):
        return self._cxx(x, y, *[theta_scope + k for k in self._hypers])

    def _layout(self):
        return [(k, dt) for k, (dt, _) in self._hypers.items()]

    @property
    def theta(self):
        return _named(self._kname, list(self._values))(**self._values)

    @theta.setter
    def theta(self, values):
        assert len(values) == len(self._values)
        for key, v in zip(self._hypers, values):
            self._values[key] = v

    @property
    def bounds(self):
        return tuple(self._bounds.values())

    @property
    def minmax(self):
        return self._minmax


class SquareExponential(_ScalarKernel):
    r""":math:`\exp(-\frac12 (x-y)^2/\ell^2)` with ``length_scale``
    :math:`\ell`."""
    _kname = 'SquareExponential'
    _hypers = {'length_scale': (np.float32, (1e-6, np.inf))}

    @staticmethod
    def _f(x, y, ls):
        d2 = (x - y) ** 2
        f = math.exp(-0.5 * d2 / ls ** 2)
        return f, [f * d2 / ls ** 3]

    @staticmethod
    def _cxx(x, y, ls):
        # exp(-d2 / (2 l^2)) = 2^(d2 * c) with c = -log2(e) / (2 l^2): the
        # parenthesised factor only depends on the hyper-parameter, so the
        # compiler hoists it out of the solver's loops and one evaluation is
        # sub, 2 mul, ex2 (the large-pair kernel evaluates the edge kernel
        # for every product of every matvec; --use_fast_math does not
        # re-associate the reference's -0.5F*d2/l^2, which costs 4 mul)
        d2 = f'graphdot::ipow<2>({x} - {y})'
        f = (f'exp2f({d2} * (-0.72134752044448170368f * '
             f'graphdot::ripow<2>({ls})))')
        return f, [f'({f} * {d2} * graphdot::ripow<3>({ls}))']


class RationalQuadratic(_ScalarKernel):
    r""":math:`(1 + (x-y)^2 / (2\alpha\ell^2))^{-\alpha}`."""
    _kname = 'RationalQuadratic'
    _hypers = {'length_scale': (np.float32, (1e-6, np.inf)),
               'alpha': (np.float32, (1e-3, np.inf))}

    @staticmethod
    def _f(x, y, ls, alpha):
        u = (x - y) ** 2 / (2 * alpha * ls ** 2)
        f = (1 + u) ** (-alpha)
        return f, [2 * alpha * u / ls * (1 + u) ** (-alpha - 1),
                   f * (u / (1 + u) - math.log1p(u))]

    @staticmethod
    def _cxx(x, y, ls, alpha):
        u = (f'(graphdot::ipow<2>({x} - {y}) * graphdot::ripow<2>({ls}) '
             f'/ (2.0f * {alpha}))')
        f = f'__powf(1.0f + {u}, -{alpha})'
        return f, [
            f'(2.0f * {alpha} * {u} / {ls} * __powf(1.0f + {u}, '
            f'-{alpha} - 1.0f))',
            f'({f} * ({u} / (1.0f + {u}) - __logf(1.0f + {u})))']


def _from_sympy(name, desc, expr, vars, *specs, minmax=(0, 1)):
    import sympy as sy
    from sympy.codegen import ast
    from sympy.printing.c import C99CodePrinter

    assert isinstance(name, str) and name.isidentifier()
    if isinstance(expr, str):
        expr = sy.sympify(expr)
    if len(vars) != 2:
        raise ValueError('A microkernel must have exactly two variables')
    vars = [sy.Symbol(v) if isinstance(v, str) else v for v in vars]

    hypers, docs = {}, {}
    for spec in specs:
        if isinstance(spec, (str, sy.Symbol)):
            spec = (spec,)
        spec = tuple(spec)
        sym = str(spec[0])
        dt = np.dtype(spec[1]) if len(spec) >= 2 else np.dtype(np.float32)
        if len(spec) in (1, 2):
            hypers[sym] = (dt, None)
        elif len(spec) == 3:
            hypers[sym], docs[sym] = (dt, None), spec[2]
        elif len(spec) == 4:
            hypers[sym] = (dt, (spec[2], spec[3]))
        elif len(spec) == 5:
            hypers[sym], docs[sym] = (dt, (spec[2], spec[3])), spec[4]
        else:
            raise ValueError('Invalid hyperparameter specification; use '
                             '(symbol[, dtype[, lb, ub][, doc]])')

    class Printer(C99CodePrinter):
        def _print_Pow(self, e):
            b, p = e.base, e.exp
            if p.is_integer and p.is_number:
                fn = 'graphdot::ipow' if int(p) >= 0 else 'graphdot::ripow'
                return f'{fn}<{abs(int(p))}>({self._print(b)})'
            return f'powf({self._print(b)}, {self._print(p)})'

    printer = Printer(dict(type_aliases={ast.real: ast.float32,
                                         ast.integer: ast.int32}))
    symbols = [*vars, *[sy.Symbol(h) for h in hypers]]
    fun = sy.lambdify(symbols, expr)
    jacs = [sy.lambdify(symbols, sy.diff(expr, sy.Symbol(h))) for h in hypers]

    def cxx(e, mapping):
        return printer.doprint(e.subs({s: sy.Symbol(m)
                                      for s, m in mapping.items()}))

    class SympyKernel(_ScalarKernel):
        _kname = name
        _hypers = hypers
        _minmax = minmax
        __doc__ = desc

        @staticmethod
        def _f(x, y, *theta):
            return fun(x, y, *theta), [j(x, y, *theta) for j in jacs]

        @staticmethod
        def _cxx(x, y, *names):
            m = {vars[0]: '_gdb_x_', vars[1]: '_gdb_y_'}
            m.update({sy.Symbol(h): f'_gdb_h{i}_'
                      for i, h in enumerate(hypers)})

            def finish(s):
                s = s.replace('_gdb_x_', x).replace('_gdb_y_', y)
                for i, n in enumerate(names):
                    s = s.replace(f'_gdb_h{i}_', n)
                return f'({s})'
            return (finish(cxx(expr, m)),
                    [finish(cxx(sy.diff(expr, sy.Symbol(h)), m))
                     for h in hypers])

    SympyKernel.__name__ = SympyKernel.__qualname__ = name
    return SympyKernel


# ----------------------------------------------------------------------------
# wrappers
# ----------------------------------------------------------------------------
class _Normalized(MicroKernel):
    def __init__(self, kernel):
        self.kernel = kernel

    @property
    def name(self):
        return 'Normalize'

    def __call__(self, x, y, jac=False):
        k = self.kernel
        if not jac:
            fxx, fxy, fyy = k(x, x), k(x, y), k(y, y)
            return fxy * (fxx * fyy) ** -0.5 if fxx > 0 and fyy > 0 else 0.0
        fxx, jxx = k(x, x, True)
        fxy, jxy = k(x, y, True)
        fyy, jyy = k(y, y, True)
        jxx, jxy, jyy = (np.asarray(j, float) for j in (jxx, jxy, jyy))
        if not (fxx > 0 and fyy > 0):
            return 0.0, np.zeros_like(jxy)
        s = fxx * fyy
        return (fxy * s ** -0.5,
                jxy * s ** -0.5 - 0.5 * fxy * s ** -1.5 * (jxx * fyy
                                                           + fxx * jyy))

    def __repr__(self):
        return f'Normalize({self.kernel!r})'

    def gen_expr(self, x, y, theta_scope=''):
        f, jac = self.kernel.gen_expr('_1', '_2', theta_scope + 'kernel.')
        fl = f'[&](auto _1, auto _2){{return {f};}}'
        return (f'normalize({fl}, {x}, {y})',
                [f'normalize_jacobian({fl}, '
                 f'[&](auto _1, auto _2){{return {j};}}, {x}, {y})'
                 for j in jac])

    def _layout(self):
        return [('kernel', self.kernel)]

    @property
    def theta(self):
        return self.kernel.theta

    @theta.setter
    def theta(self, values):
        self.kernel.theta = values

    @property
    def bounds(self):
        return self.kernel.bounds

    @property
    def minmax(self):
        lo, hi = self.kernel.minmax
        return (lo / hi, 1)


def Normalize(kernel):
    r""":math:`k(x,y)/\sqrt{k(x,x)k(y,y)}`; idempotent."""
    return kernel if kernel.name == 'Normalize' else _Normalized(kernel)


class _CompositeKernel(MicroKernel):
    _opnames = {'+': 'Additive', '*': 'Product'}

    def __init__(self, opstr, **kw_kernels):
        if opstr not in self._opnames:
            raise ValueError(f'Invalid reduction operator {opstr!r}.')
        self.opstr = opstr
        self.kw_kernels = kw_kernels

    def __getattr__(self, key):
        kws = self.__dict__.get('kw_kernels')
        if kws is not None and key in kws:
            return kws[key]
        raise AttributeError(key)

    @property
    def name(self):
        return 'Composite'

    @property
    def opname(self):
        return self._opnames[self.opstr]

    def __repr__(self):
        inner = ', '.join(f'{k}={v!r}' for k, v in self.kw_kernels.items())
        return f'Composite({self.opstr!r}, {inner})'

    def __call__(self, x, y, jac=False):
        if not jac:
            vals = [k(x[key], y[key]) for key, k in self.kw_kernels.items()]
            return sum(vals) if self.opstr == '+' else math.prod(vals)
        fs, js = [], []
        for key, k in self.kw_kernels.items():
            f, j = k(x[key], y[key], True)
            fs.append(f)
            js.append(np.asarray(j, float).ravel())
        if self.opstr == '+':
            return sum(fs), np.concatenate(js) if js else np.array([])
        total = math.prod(fs)
        parts = []
        for i, j in enumerate(js):
            others = math.prod(fs[:i] + fs[i + 1:])
            parts.append(others * j)
        return total, np.concatenate(parts) if parts else np.array([])

    def gen_expr(self, x, y, theta_scope=''):
        fs, js = [], []
        for key, k in self.kw_kernels.items():
            f, j = k.gen_expr(f'{x}.{key}', f'{y}.{key}',
                              f'{theta_scope}{key}.')
            fs.append(f)
            js.append(j)
        value = '(' + f' {self.opstr} '.join(fs) + ')'
        jac = []
        for i, j_list in enumerate(js):
            for j in j_list:
                if self.opstr == '+':
                    jac.append(j)
                else:
                    jac.append('(' + ' * '.join(fs[:i] + [j] + fs[i + 1:])
                               + ')')
        return value, jac

    def _layout(self):
        return list(self.kw_kernels.items())

    @property
    def theta(self):
        return _named('Composite', self.kw_kernels)(
            *[k.theta for k in self.kw_kernels.values()])

    @theta.setter
    def theta(self, values):
        for k, v in zip(self.kw_kernels.values(), values):
            k.theta = v

    @property
    def bounds(self):
        return _named('Composite', self.kw_kernels)(
            *[k.bounds for k in self.kw_kernels.values()])

    @property
    def minmax(self):
        mm = np.array([k.minmax for k in self.kw_kernels.values()])
        return (mm.sum(axis=0) if self.opstr == '+' else mm.prod(axis=0))


def Composite(oper, **kw_kernels):
    r"""Combine per-feature kernels with ``'+'`` or ``'*'``:
    :math:`k(X,Y)=k_{a_1}(X_{a_1},Y_{a_1})\,\mathrm{op}\,k_{a_2}(\ldots)`."""
    return _CompositeKernel(oper, **kw_kernels)


def TensorProduct(**kw_kernels):
    return Composite('*', **kw_kernels)


def Additive(**kw_kernels):
    return Composite('+', **kw_kernels)


class _ConvolutionKernel(MicroKernel):
    def __init__(self, kernel, mean=True):
        self.kernel = kernel
        self.mean = mean

    @property
    def name(self):
        return 'Convolution'

    def __call__(self, x, y, jac=False):
        reduce = np.mean if self.mean else np.sum
        if not jac:
            return reduce([self.kernel(a, b) for a in x for b in y])
        fs, js = zip(*[self.kernel(a, b, True) for a in x for b in y])
        return reduce(fs), reduce(np.asarray(js, float), axis=0)

    def __repr__(self):
        return f'Convolution({self.kernel!r})' if self.mean else \
            f'Convolution({self.kernel!r}, mean=False)'

    def gen_expr(self, x, y, theta_scope=''):
        f, jac = self.kernel.gen_expr('_1', '_2', theta_scope + 'kernel.')
        m = 'true' if self.mean else 'false'
        return (f'convolution<{m}>([&](auto _1, auto _2){{return {f};}}, '
                f'{x}, {y})',
                [f'convolution_jacobian<{m}>([&](auto _1, auto _2)'
                 f'{{return {j};}}, {x}, {y})' for j in jac])

    def _layout(self):
        return [('kernel', self.kernel)]

    @property
    def theta(self):
        return _named('Convolution', ['base'])(self.kernel.theta)

    @theta.setter
    def theta(self, values):
        self.kernel.theta = values[0]

    @property
    def bounds(self):
        return (self.kernel.bounds,)

    @property
    def minmax(self):
        return self.kernel.minmax


def Convolution(kernel, mean=True):
    r"""Average (``mean=True``) or sum of a base kernel over all pairs of
    elements of two variable-length feature sequences."""
    return _ConvolutionKernel(kernel, mean)
