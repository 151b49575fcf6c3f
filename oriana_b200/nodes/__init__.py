"""Graph nodes (oriana/nodes/): the API shell around the device expectations."""
from .base import Node, DeterministicNode, ProbabilisticNode
from .deterministic import Einsum, Multiply, Transpose
from .probabilistic import Gamma, Bernoulli, Poisson, Multinomial

__all__ = ['Node', 'DeterministicNode', 'ProbabilisticNode', 'Einsum', 'Multiply', 'Transpose',
           'Gamma', 'Bernoulli', 'Poisson', 'Multinomial']
