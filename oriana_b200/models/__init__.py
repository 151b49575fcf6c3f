"""Factor models (oriana/models/__init__.py).  ZIGaP and GaP are the models of the accelerated CAVI path (tensor
kernels); SparseZIGaP runs the same iteration with the sparsity layer on the CUDA-core
kernels.  SparseGaP.step() raises NameError in the reference (sparse_gap.py:127): the name resolves, constructing it explains."""
from .base import FactorModel
from .gap import GaP
from .zigap import ZIGaP
from .sparse_zigap import SparseZIGaP
from .sparse_gap import SparseGaP

__all__ = ['FactorModel', 'GaP', 'ZIGaP', 'SparseZIGaP', 'SparseGaP']
