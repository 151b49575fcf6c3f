"""oriana_b200 -- B200-native (sm_100a) implementation of Oriana's PCMF CAVI iteration.

Mirrors the import surface of the reference package (`oriana/__init__.py:1-3`):
`Dimensions`, `DimRelation`, `Parameter`, the two exception types; sub-packages `models`, `nodes`,
`inference`, `utils`, `singlecell`.  All arithmetic of `step()` runs in hand-written CUDA kernels
behind the C ABI of `include/oriana_b200.h`; there is no CPU fallback.
"""
from .exceptions import DatatypeException, IncompatibleShapeException
from .parameters import Parameter
from .dims import Dimensions, DimRelation

__all__ = ['DatatypeException', 'IncompatibleShapeException', 'Parameter', 'Dimensions', 'DimRelation']
__version__ = '0.1.0'
