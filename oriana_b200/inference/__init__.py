from .variational import VariationalDistribution

__all__ = ['VariationalDistribution']
