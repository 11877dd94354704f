"""Preset kernel for 3D molecular graphs (reference
graphdot/kernel/molecular.py:13-91; Tang & de Jong, J. Chem. Phys. 150,
044107 (2019))."""
import copy

from ..microkernel import KroneckerDelta, SquareExponential, TensorProduct
from .marginalized import MarginalizedGraphKernel


class Tang2019MolecularKernel:
    def __init__(self, stopping_probability=0.01, starting_probability=1.0,
                 element_prior=0.2, edge_length_scale=0.05, **kwargs):
        self.stopping_probability = stopping_probability
        self.starting_probability = starting_probability
        self.element_prior = element_prior
        self.edge_length_scale = edge_length_scale
        self.kernel = MarginalizedGraphKernel(
            TensorProduct(element=KroneckerDelta(element_prior)),
            TensorProduct(length=SquareExponential(edge_length_scale)),
            q=stopping_probability, p=starting_probability, **kwargs)

    def __call__(self, X, Y=None, **kwargs):
        return self.kernel(X, Y, **kwargs)

    def diag(self, X, **kwargs):
        return self.kernel.diag(X, **kwargs)

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
