"""Deterministic nodes (oriana/nodes/deterministic/{einsum,multiply,transpose}.py).

Out of the hot path: `step()` never evaluates them (the reference materialises U V^T as an n x p float64
array every step, zigap.py:141; here `UV.forward()` is only run on request).  Plain torch ops on the
device the parents live on.
"""
import torch

from .base import DeterministicNode


class Einsum(DeterministicNode):

    def __init__(self, subscripts, *nodes, **kwargs):
        DeterministicNode.__init__(self, *nodes, **kwargs)
        self.subscripts = subscripts

    def _sample(self, *params):
        return torch.einsum(self.subscripts, *params)


class Multiply(DeterministicNode):

    def __init__(self, left_node, right_node, **kwargs):
        DeterministicNode.__init__(self, left_node, right_node, **kwargs)

    def _sample(self, left, right):
        return left * right


class Transpose(DeterministicNode):

    def __init__(self, node, **kwargs):
        DeterministicNode.__init__(self, node, **kwargs)

    def _sample(self, arr):
        return arr.T
