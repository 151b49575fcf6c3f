"""Factor models (oriana/models/__init__.py).  ZIGaP and GaP are the models of the accelerated CAVI path;
the sparse variants (sparse_zigap.py, sparse_gap.py) are the next row of SURVEY.md section 8f."""
from .base import FactorModel
from .gap import GaP
from .zigap import ZIGaP

__all__ = ['FactorModel', 'GaP', 'ZIGaP']
