"""Front end of the marginalized graph kernel (MLGK).

Same public surface as the reference's ``MarginalizedGraphKernel`` (reference
graphdot/kernel/marginalized/_kernel.py:17-508): construction arguments,
``__call__(X, Y, eval_gradient, nodal, lmin)``, ``diag``, and the
scikit-learn style ``theta / bounds / hyperparameters / clone_with_theta``.
The front end only lays out jobs and outputs; all arithmetic happens in the
back end (``B200Backend`` -> libgraphdot_b200.so -> sm_100a kernels).

Output conventions kept from the reference: the back end writes float32 in
Fortran order, ``gramian[r + c*nX]`` and ``gradient[r + c*nX + k*nX*nY]``
(reference _kernel.py:247-251, graphdot/cpp/tensor_view.h:24-33); the Jacobian
covers *all* hyper-parameters in the order ``[p..., q, node..., edge...]`` and
fixed ones are masked on the host; derivatives are w.r.t. the hyper-parameters
themselves, not their logarithms (reference model/gaussian_process/gpr.py:296
applies the chain rule).
"""
import copy
import itertools as it
import os
import numbers
import warnings
from collections import namedtuple

import numpy as np

from ...graph import Graph, VolatileCookie
from ...util import Timer, flatten, fold_like, replace
from ._backend import backend_factory
from .starting_probability import Adhoc, StartingProbability, Uniform

_Hyper = namedtuple('MarginalizedGraphKernel',
                    ['starting_probability', 'stopping_probability',
                     'node_kernel', 'edge_kernel'])
_HyperBounds = namedtuple('GraphKernelHyperparameterBounds', _Hyper._fields)
JOB_DTYPE = np.dtype([('i', np.uint32), ('j', np.uint32)])


class MarginalizedGraphKernel:
    """Random-walk graph similarity kernel of Kashima, Tsuda & Inokuchi
    (ICML 2003) in the generalized-Laplacian formulation of Tang & de Jong
    (J. Chem. Phys. 150, 044107).

    Parameters
    ----------
    node_kernel, edge_kernel: microkernels
        Similarity between individual nodes / edges.
    p: positive number or StartingProbability or (callable, C++ expr)
        Starting probability of the random walk on each node.
    q: float in (0, 1)
        Stopping probability; ``q_bounds`` its range during training.
    eps: float
        Step (in log-hyper-parameter) of the finite differences used for
        nodal gradients.
    ftol, gtol: float
        CG tolerances of the value solve / warm-started re-solves.
    dtype: numpy dtype of the returned matrices.
    backend: 'auto' | 'cuda' | 'b200' | Backend instance.
    """
    trait_t = namedtuple('Traits',
                         'diagonal, symmetric, nodal, lmin, eval_gradient')

    @classmethod
    def traits(cls, diagonal=False, symmetric=False, nodal=False, lmin=0,
               eval_gradient=False):
        return cls.trait_t(diagonal, symmetric, nodal, lmin, eval_gradient)

    def __init__(self, node_kernel, edge_kernel, p=1.0, q=0.01,
                 q_bounds=(1e-4, 1 - 1e-4), eps=1e-2, ftol=1e-8, gtol=1e-6,
                 dtype=float, backend='auto'):
        self.node_kernel = node_kernel
        self.edge_kernel = edge_kernel
        self.p = self._get_starting_probability(p)
        self.q = q
        self.q_bounds = q_bounds
        self.eps = eps
        self.ftol = ftol
        self.gtol = gtol
        self.element_dtype = dtype
        self.backend = backend_factory(backend)

        lo, hi = self.node_kernel.minmax
        if lo <= 0 or hi > 1:
            warnings.warn(
                'Node kernel value range should be within (0, 1], got '
                f'{self.node_kernel.minmax} for {self.node_kernel}. Consider '
                'adding a small constant or using `.normalized`.',
                DeprecationWarning)
        lo, hi = self.edge_kernel.minmax
        if lo < 0 or hi > 1:
            warnings.warn(
                'Edge kernel value range must be within [0, 1], got '
                f'{self.edge_kernel.minmax} for {self.edge_kernel}. Consider '
                'adding a small constant or using `.normalized`.',
                DeprecationWarning)

    @staticmethod
    def _get_starting_probability(p):
        if isinstance(p, StartingProbability):
            return p
        if isinstance(p, tuple) and len(p) == 2:
            f, expr = p
            if callable(f) and isinstance(expr, str):
                return Adhoc(f, expr)
            raise ValueError('An ad hoc starting probability must be a '
                             '(callable, C++ expression) pair.')
        if isinstance(p, numbers.Number):
            if p > 0:
                return Uniform(p)
            raise ValueError(f'Starting probability {p} < 0.')
        raise ValueError(f'Unknown starting probability: {p}')

    # ------------------------------------------------------------------
    @staticmethod
    def _check_types(graphs):
        verdict = Graph.has_unified_types(graphs)
        if verdict is not True:
            group, first, second = verdict
            raise TypeError(
                f'The two graphs have mismatching {group} attributes or '
                'attribute types. If the attributes match in name but differ '
                'in type, try `Graph.unify_datatype` as an automatic fix.\n'
                f'First graph: {first}\nSecond graph: {second}\n')

    def normalized_gram(self, X, Y=None, eval_gradient=False, lmin=0,
                        timing=False, **unsupported):
        """K_ij / sqrt(K_ii K_jj) (and its Jacobian) with the normalization
        fused into the solver's epilogue: one extra launch for the
        self-similarities, no host post-processing.  Returns NotImplemented
        when the back end cannot do it (``Normalization`` then falls back to
        the reference's host formulas, reference kernel/fix.py:46-73)."""
        if unsupported or not getattr(self.backend, 'fused_normalization',
                                      False):
            return NotImplemented
        return self.__call__(X, Y, eval_gradient=eval_gradient, lmin=lmin,
                             timing=timing, _fused_normalization=True)

    def device_gram(self, X, Y=None, eval_gradient=False, lmin=0,
                    normalize=False):
        """Graph-level Gram matrix (and Jacobian) as float32 torch CUDA
        tensors that never visit the host -- for callers that continue on the
        device, e.g. ``graphdot_b200.model.gaussian_process``.  ``normalize``
        fuses K_ij / sqrt(K_ii K_jj) into the solve."""
        if not hasattr(self.backend, 'device_outputs'):
            raise NotImplementedError('back end has no device-resident outputs')
        return self.__call__(X, Y, eval_gradient=eval_gradient, lmin=lmin,
                             _fused_normalization=bool(normalize),
                             _device=True)

    def __call__(self, X, Y=None, eval_gradient=False, nodal=False, lmin=0,
                 timing=False, _fused_normalization=False, _device=False):
        """Pairwise similarity matrix between the graphs in ``X`` (and ``Y``).

        Returns the (len(X), len(Y)) matrix -- node-by-node if ``nodal`` --
        and, with ``eval_gradient``, its derivative with respect to every
        non-fixed hyper-parameter stacked along a third axis.  ``lmin=1``
        drops the zero-length paths from the similarity."""
        timer = Timer()
        backend = self.backend
        traits = self.traits(symmetric=Y is None, nodal=nodal, lmin=lmin,
                             eval_gradient=eval_gradient)
        graphs = list(X) if Y is None else list(it.chain(X, Y))
        # the type check walks every graph (2 ms for 2000): once per set of
        # graph objects, not once per call (training loops)
        # -- and not again until a graph's cache is invalidated (in-place
        # permutation, unify_datatype)
        ids = tuple(map(id, graphs))
        checked = getattr(self, '_types_checked', None)
        if (checked is None or checked[0] != ids
                or checked[2] != VolatileCookie.epoch):
            self._check_types(graphs)
            own = all(isinstance(g, Graph) for g in graphs)
            self._types_checked = ((ids, graphs, VolatileCookie.epoch)
                                   if own else None)
        nx, ny = len(X), (len(X) if Y is None else len(Y))

        timer.tic('generating jobs')
        grid = getattr(backend, 'pair_jobs', None)
        if grid is not None:
            # implicit job grid, decoded on the device
            pairs = (grid.triu(0, nx) if traits.symmetric
                     else grid.rect(0, nx, nx, nx + ny))
        else:
            if traits.symmetric:
                i, j = np.triu_indices(nx)
            else:
                i, j = np.divmod(np.arange(nx * ny), ny)
                j = j + nx
            pairs = np.empty(len(i), dtype=JOB_DTYPE)
            pairs['i'], pairs['j'] = i, j
        jobs = backend.array(pairs)
        timer.toc('generating jobs')

        timer.tic('creating output buffer')
        starts = backend.zeros(len(graphs) + 1, dtype=np.uint32)
        if traits.nodal is True:
            sizes = np.array([len(g.nodes) for g in graphs], dtype=np.uint32)
            starts[1:nx + 1] = np.cumsum(sizes[:nx])
            rows = cols = int(starts[nx])
            if not traits.symmetric:
                starts[nx] = 0
                starts[nx + 1:] = np.cumsum(sizes[nx:])
                cols = int(starts[-1])
        else:
            starts[:nx] = np.arange(nx)
            if traits.symmetric:
                starts[nx] = nx
            else:
                starts[nx:] = np.arange(ny + 1)
            rows, cols = nx, ny
        want_grad = traits.eval_gradient is True
        coll = None
        if _device:
            # results go straight into torch-owned device tensors: C order
            # (k, c, r) is the solver's Fortran order [r + c rows + k rows cols]
            import torch
            dev = torch.device('cuda', backend.device)
            Kt = torch.zeros((cols, rows), dtype=torch.float32, device=dev)
            dKt = (torch.zeros((self.n_dims, cols, rows), dtype=torch.float32,
                               device=dev) if want_grad else None)
            gramian = gradient = None
            extra_out = dict(
                gramian_dev=Kt.data_ptr(),
                gradient_dev=dKt.data_ptr() if want_grad else None,
                stream=torch.cuda.current_stream(dev).cuda_stream)
        else:
            gramian = backend.empty(rows * cols, np.float32)
            gradient = (backend.empty(self.n_dims * rows * cols, np.float32)
                        if want_grad else None)
            extra_out = {}
            make = getattr(backend, 'collect', None)
            mask = np.asarray(self.active_theta_mask, dtype=bool)
            dtype = np.dtype(self.element_dtype)
            plain = dtype == np.float32 and (not want_grad or mask.all())
            if (make is not None and not plain
                    and dtype in (np.float32, np.float64)):
                # conversion to `dtype` and the active-plane selection happen
                # per finished column block on the engine's host threads
                coll = make(rows, cols, self.n_dims if want_grad else 0,
                            mask, dtype)
                extra_out['collect'] = coll
            if make is not None and len(pairs) >= 65536:
                # pipeline: a handful of launches, copy-back and collection
                # of a finished column block overlap the next launch.
                # (C3 on a B200, end to end: 4 uniform launches 103.9 ms, 8:
                # 105.2, 12: 106.7; shrinking launch sizes did not help.)
                n_launch = int(os.environ.get('GDB_PIPELINE_LAUNCHES', 4))
                extra_out['tile'] = max(32, -(-(nx if traits.symmetric
                                                else ny) // n_launch))
        timer.toc('creating output buffer')

        timer.tic('calling GPU kernel (overall)')
        extra = dict(extra_out)
        if _fused_normalization:
            # self-similarities of every graph stay on the device ...
            n = len(graphs)
            djobs = np.empty(n, dtype=JOB_DTYPE)
            djobs['i'] = djobs['j'] = np.arange(n)
            backend(graphs, self.node_kernel, self.edge_kernel, self.p,
                    self.q, self.eps, self.ftol, self.gtol,
                    backend.array(djobs),
                    np.arange(n + 1, dtype=np.uint32),
                    None, None, n, 1, self.n_dims,
                    self.traits(diagonal=True, lmin=lmin,
                                eval_gradient=eval_gradient),
                    timer, store_diag=True, keep_on_device=True,
                    stream=extra.get('stream'))
            extra['normalize'] = True   # ... and scale the main solve
            extra['resend'] = False     # the graphs were (re-)sent just now
        backend(graphs, self.node_kernel, self.edge_kernel, self.p, self.q,
                self.eps, self.ftol, self.gtol, jobs, starts, gramian,
                gradient, rows, cols, self.n_dims, traits, timer, **extra)
        timer.toc('calling GPU kernel (overall)')

        if _device:
            K = Kt.t()
            if dKt is None:
                return K
            mask = torch.as_tensor(np.asarray(self.active_theta_mask),
                                   device=dev)
            return K, dKt.permute(2, 1, 0)[:, :, mask]

        timer.tic('collecting result')
        if coll is not None:
            gramian, gradient = coll.gram, coll.grad
        else:
            gramian = gramian.reshape(rows, cols, order='F')
            if gradient is not None:
                gradient = self._active_planes(
                    gradient.reshape((rows, cols, self.n_dims), order='F'),
                    self.active_theta_mask, self.element_dtype)
            gramian = gramian.astype(self.element_dtype, copy=False)
        timer.toc('collecting result')
        if timing:
            timer.report(unit='ms')

        if gradient is not None:
            return gramian, gradient
        return gramian

    @staticmethod
    def _active_planes(jacobian, mask, dtype):
        """``jacobian[:, :, mask].astype(dtype)`` in ONE pass over the data
        (the Jacobian of 2000 graphs is 80 MB in float32): a plain conversion
        when every hyper-parameter is active, otherwise plane-by-plane copies
        into the Fortran-ordered result."""
        mask = np.asarray(mask, dtype=bool)
        if mask.all():
            return jacobian.astype(dtype)
        out = np.empty(jacobian.shape[:2] + (int(mask.sum()),), dtype=dtype,
                       order='F')
        for k, plane in enumerate(np.flatnonzero(mask)):
            out[:, :, k] = jacobian[:, :, plane]
        return out

    def diag(self, X, eval_gradient=False, nodal=False, lmin=0,
             active_theta_only=True, timing=False):
        """Self-similarities of the graphs in ``X``: a vector of graph
        (``nodal=False``) or node (``nodal=True``) self-similarities, or with
        ``nodal='block'`` the list of per-graph nodal similarity matrices."""
        timer = Timer()
        backend = self.backend
        if nodal not in (True, False, 'block'):
            raise ValueError("Invalid 'nodal' option '%s'" % nodal)
        traits = self.traits(diagonal=True, nodal=nodal, lmin=lmin,
                             eval_gradient=eval_gradient)
        X = list(X)
        self._check_types(X)

        timer.tic('generating jobs')
        pairs = np.empty(len(X), dtype=JOB_DTYPE)
        pairs['i'] = pairs['j'] = np.arange(len(X))
        jobs = backend.array(pairs)
        timer.toc('generating jobs')

        timer.tic('creating output buffer')
        starts = backend.zeros(len(X) + 1, dtype=np.uint32)
        sizes = np.array([len(g.nodes) for g in X], dtype=np.uint32)
        if nodal is True:
            starts[1:] = np.cumsum(sizes)
        elif nodal is False:
            starts[:] = np.arange(len(X) + 1)
        else:
            starts[1:] = np.cumsum(sizes.astype(np.uint64) ** 2)
        length = int(starts[-1])
        gramian = backend.empty(length, np.float32)
        gradient = (backend.empty(self.n_dims * length, np.float32)
                    if traits.eval_gradient is True else None)
        timer.toc('creating output buffer')

        timer.tic('calling GPU kernel (overall)')
        backend(X, self.node_kernel, self.edge_kernel, self.p, self.q,
                self.eps, self.ftol, self.gtol, jobs, starts, gramian,
                gradient, length, 1, self.n_dims, traits, timer)
        timer.toc('calling GPU kernel (overall)')

        timer.tic('collecting result')
        if gradient is not None:
            gradient = gradient.reshape((length, self.n_dims), order='F')
            if active_theta_only:
                gradient = gradient[:, self.active_theta_mask]
        if nodal == 'block':
            out = [np.array(gramian[s:s + n * n]).reshape(n, n)
                   for s, n in zip(starts[:-1], sizes.astype(int))]
        elif gradient is not None:
            out = (gramian.astype(self.element_dtype),
                   gradient.astype(self.element_dtype))
        else:
            out = gramian.astype(self.element_dtype)
        timer.toc('collecting result')
        if timing:
            timer.report(unit='ms')
        return out

    # ---- scikit-learn interoperability --------------------------------
    def is_stationary(self):
        return False

    @property
    def requires_vector_input(self):
        return False

    @property
    def hyperparameters(self):
        """Hierarchical view of all hyper-parameters."""
        return _Hyper(self.p.theta, self.q, self.node_kernel.theta,
                      self.edge_kernel.theta)

    @property
    def flat_hyperparameters(self):
        return np.fromiter(flatten(self.hyperparameters), float)

    @property
    def hyperparameter_bounds(self):
        return _HyperBounds(self.p.bounds, self.q_bounds,
                            self.node_kernel.bounds, self.edge_kernel.bounds)

    def _flat_bounds(self):
        flat = flatten(replace(flatten(self.hyperparameter_bounds), 'fixed',
                               (np.nan, np.nan)))
        return np.fromiter(flat, float).reshape(-1, 2)

    @property
    def n_dims(self):
        """Number of hyper-parameters, optimizable and fixed."""
        return len(self.flat_hyperparameters)

    @property
    def active_theta_mask(self):
        lower, upper = self._flat_bounds().T
        return ~(np.isnan(lower) | np.isnan(upper) | (lower == upper))

    @property
    def theta(self):
        """Logarithms of the non-fixed hyper-parameters, flattened."""
        return np.log(self.flat_hyperparameters[self.active_theta_mask])

    @theta.setter
    def theta(self, value):
        logs = np.log(self.flat_hyperparameters)
        logs[self.active_theta_mask] = value
        (self.p.theta, self.q, self.node_kernel.theta,
         self.edge_kernel.theta) = fold_like(np.exp(logs),
                                             self.hyperparameters)

    @property
    def bounds(self):
        """Log-bounds of the non-fixed hyper-parameters, shape (n, 2)."""
        return np.log(self._flat_bounds()[self.active_theta_mask, :])

    def clone_with_theta(self, theta):
        checked, self._types_checked = getattr(self, '_types_checked', None), None
        try:
            clone = copy.deepcopy(self)      # without the memo of checked graphs
        finally:
            self._types_checked = checked
        clone._types_checked = checked       # same graphs, same verdict
        clone.theta = theta
        return clone
