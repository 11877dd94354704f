"""Inputs shared by ``make_gpr_golden.py`` (reference side, build container)
and ``tests/test_gpr.py`` (this package): a numpy kernel with two log-scale
hyper-parameters and a fixed data set with one masked target.

TEST INFRASTRUCTURE."""
import numpy as np


class RBF:
    """s^2 exp(-d^2 / 2 L^2) on scalars, Jacobian with respect to (s, L)."""

    def __init__(self, s, L):
        self.s, self.L = s, L

    def __call__(self, X, Y=None, eval_gradient=False):
        d = np.subtract.outer(X, Y if Y is not None else X)
        e = np.exp(-0.5 * d ** 2 / self.L ** 2)
        f = self.s ** 2 * e
        if eval_gradient is False:
            return f
        return f, np.stack((2 * self.s * e, f * d ** 2 * self.L ** -3), axis=2)

    def diag(self, X):
        return np.full(len(X), self.s ** 2)

    @property
    def theta(self):
        return np.log([self.s, self.L])

    @theta.setter
    def theta(self, t):
        self.s, self.L = np.exp(t)

    @property
    def bounds(self):
        return np.log([[1e-2, 1e2], [1e-2, 1e2]])

    def clone_with_theta(self, theta):
        k = RBF(1.0, 1.0)
        k.theta = theta
        return k


def data():
    rng = np.random.default_rng(42)
    X = np.sort(rng.uniform(-2, 2, 14))
    y = np.sin(1.7 * X) + 0.3 * X + 0.05 * rng.standard_normal(14)
    y_masked = y.copy()
    y_masked[5] = np.nan
    Z = np.linspace(-2.2, 2.2, 9)
    return X, y, y_masked, Z


THETAS = [(0.7, 0.4), (1.0, 1.0), (1.8, 0.25)]
