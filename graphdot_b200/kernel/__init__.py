from .marginalized import MarginalizedGraphKernel
from .molecular import Tang2019MolecularKernel
from .fix import Normalization

__all__ = ['MarginalizedGraphKernel', 'Tang2019MolecularKernel',
           'Normalization']
