"""Kernel decorators (reference graphdot/kernel/fix.py:8-117).

``Normalization(kernel)`` returns K_ij / sqrt(K_ii K_jj) and, on request, its
gradient by the quotient rule.  For symmetric calls the diagonal is read from
the Gram matrix itself, for X-by-Y calls from ``kernel.diag`` (reference
fix.py:36-41, :64-69)."""
import copy

import numpy as np


class Normalization:
    def __init__(self, kernel):
        self.kernel = kernel

    def __call__(self, X, Y=None, eval_gradient=False, **options):
        fused = getattr(self.kernel, 'normalized_gram', None)
        if fused is not None and not options.get('nodal', False):
            out = fused(X, Y, eval_gradient=eval_gradient, **options)
            if out is not NotImplemented:
                return out
        return self._host_normalized(X, Y, eval_gradient, **options)

    def device_gram(self, X, Y=None, eval_gradient=False, **options):
        """Normalized Gram matrix (and Jacobian) as device-resident torch
        tensors (see ``MarginalizedGraphKernel.device_gram``)."""
        return self.kernel.device_gram(X, Y, eval_gradient=eval_gradient,
                                       normalize=True, **options)

    def _host_normalized(self, X, Y=None, eval_gradient=False, **options):
        if eval_gradient is True:
            R, dR = self.kernel(X, Y, eval_gradient=True, **options)
            if Y is None:
                dl = dr = R.diagonal()
                ddl = ddr = np.einsum('iik->ik', dR)
            else:
                dl, ddl = self.kernel.diag(X, True, **options)
                dr, ddr = self.kernel.diag(Y, True, **options)
        else:
            R = self.kernel(X, Y, **options)
            if Y is None:
                dl = dr = R.diagonal()
            else:
                dl = self.kernel.diag(X, **options)
                dr = self.kernel.diag(Y, **options)
        sl, sr = dl ** -0.5, dr ** -0.5
        K = sl[:, None] * R * sr[None, :]
        if eval_gradient is not True:
            return K
        # d(R/sqrt(a b)) = dR/sqrt(ab) - K/2 (da/a + db/b)
        dK = (sl[:, None, None] * dR * sr[None, :, None]
              - 0.5 * K[:, :, None] * ((ddl / dl[:, None])[:, None, :]
                                       + (ddr / dr[:, None])[None, :, :]))
        return K, np.asfortranarray(dK)

    def diag(self, X, eval_gradient=False, **options):
        if eval_gradient is True:
            return np.ones(len(X)), np.ones((len(X), len(self.kernel.theta)))
        return np.ones(len(X))

    hyperparameters = property(lambda self: self.kernel.hyperparameters)
    hyperparameter_bounds = property(
        lambda self: self.kernel.hyperparameter_bounds)
    bounds = property(lambda self: self.kernel.bounds)

    @property
    def theta(self):
        return self.kernel.theta

    @theta.setter
    def theta(self, value):
        self.kernel.theta = value

    def clone_with_theta(self, theta):
        clone = copy.deepcopy(self)
        clone.theta = theta
        return clone
