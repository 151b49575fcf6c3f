"""Node base classes (oriana/nodes/base.py:10-172).

A node owns a buffer (a torch tensor, in HBM when a GPU is present) and knows its parents.  The
reference funnels `mean()/meanlog()/sample()` through the `updates_buffer` decorator
(nodes/base.py:123-143): gather the parents' arrays, evaluate, check the canonical (s, d, c) shape, fold it
back to the node's shape, store it in the buffer and return it.  `_evaluate` below is that funnel.
Expectations are computed by CUDA kernels (see probabilistic.py); results are returned as host numpy
arrays, like the reference, and kept on the device in `node.tensor`.
"""
import numpy as np
import torch

from ..parameters import Parameter, as_tensor, default_device


def _parent_tensor(parent):
    if isinstance(parent, (Node, Parameter)):
        return parent.tensor
    return as_tensor(parent)


class Node:

    def __init__(self, *parents):
        self.parents = parents
        self.children = []
        self.fixed = False
        self._t = None
        self.integer_samples = False
        for parent in parents:
            if isinstance(parent, Node):
                parent.add_child(self)

    def add_child(self, node):
        if node not in self.children:
            self.children.append(node)

    def fix(self, recursive=False):
        self._set_fixed(True, recursive)

    def unfix(self, recursive=False):
        self._set_fixed(False, recursive)

    def _set_fixed(self, flag, recursive):
        self.fixed = flag
        if recursive:
            for parent in self.parents:
                if isinstance(parent, Node):
                    parent._set_fixed(flag, True)

    # -- buffer access (nodes/base.py:42-61) ---------------------------------------------------------
    @property
    def tensor(self):
        # buffers of probabilistic nodes are allocated on first use: a model-level D node is n x p
        if self._t is None and getattr(self, '_lazy_shape', None) is not None:
            self._t = torch.zeros(self._lazy_shape, dtype=torch.float64, device=default_device())
        return self._t

    @property
    def buffer(self):
        t = self.tensor
        return None if t is None else t.detach().cpu().numpy()

    @buffer.setter
    def buffer(self, data):
        # `integer_samples`: the reference keeps the dtype of whatever was assigned, and Poisson.logp() allocates its
        # result with it (poisson.py:68) -- per-entry log-probabilities of integer counts are truncated to integers
        if isinstance(data, torch.Tensor):
            self._t = data                       # adopt a device tensor as is (no copy, any dtype)
            self.integer_samples = not (data.dtype.is_floating_point or data.dtype.is_complex)
        else:
            self.integer_samples = bool(np.issubdtype(np.asarray(data).dtype, np.integer))
            dtype = self._t.dtype if self._t is not None else torch.float64
            self._t = as_tensor(data, dtype=dtype)

    def asarray(self):
        return np.asarray(self.buffer)

    def __getitem__(self, key):
        return self.buffer[key]

    def __setitem__(self, key, value):
        if isinstance(value, (Node, Parameter)):
            value = value.tensor
        if not isinstance(value, torch.Tensor):
            value = torch.as_tensor(np.asarray(value))
        t = self.tensor
        if isinstance(key, np.ndarray):
            key = torch.as_tensor(key, device=t.device)
        t[key] = value.to(device=t.device, dtype=t.dtype)

    def sample(self, **kwargs):
        raise NotImplementedError


class DeterministicNode(Node):
    """Buffer = f(parents), recomputed only when `forward()`/`sample()` is called (nodes/base.py:64-90)."""

    def __init__(self, *parents, name=''):
        Node.__init__(self, *parents)
        self.name = name

    def sample(self, recursive=False):
        args = []
        for parent in self.parents:
            if isinstance(parent, Node) and recursive:
                parent.sample(recursive=recursive)
            args.append(_parent_tensor(parent))
        if not self.fixed:
            self._t = self._sample(*args)
        assert self._t is not None
        return self.buffer

    def forward(self):
        return self.sample(recursive=False)

    def _sample(self, *params):
        raise NotImplementedError


class ProbabilisticNode(Node):
    """A random variable with a dimension relation (nodes/base.py:93-172)."""

    def __init__(self, *parents, rel=None, name=''):
        Node.__init__(self, *parents)
        self.name = name
        self.rel = rel
        self.shape = rel.shape
        self.n_samples_per_distrib = rel.n_samples_per_distrib
        self.n_distribs = rel.n_distribs
        self.n_components = rel.n_components
        self.reshape_func = rel.reshape_func
        self.inv_reshape_func = rel.inv_reshape_func
        self._lazy_shape = self.shape

    def _evaluate(self, func, recursive=False):
        args = []
        for parent in self.parents:
            if isinstance(parent, Node) and recursive:
                parent.sample(recursive=recursive)
            args.append(_parent_tensor(parent))
        if self.fixed:
            return self.buffer
        out = func(*args)
        assert tuple(out.shape) == self.rel.canonical_shape
        out = self.reshape_func(out)
        self.tensor[...] = out.to(self.tensor.dtype)
        # the returned array keeps the dtype of the evaluation (float32 for meanlog), like the reference
        return out.detach().cpu().numpy()

    def sample(self, recursive=False):
        return self._evaluate(self._sample, recursive)

    def mean(self, recursive=False):
        return self._evaluate(self._mean, recursive)

    def logp(self):
        samples = self.inv_reshape_func(self.tensor)
        args = [_parent_tensor(parent) for parent in self.parents]
        return torch.nan_to_num(self._logp(samples, *args)).cpu().numpy()

    def loglikelihood(self):
        return float(self.logp().sum())

    def _sample(self, *params):
        raise NotImplementedError

    def _mean(self, *params):
        raise NotImplementedError

    def _logp(self, samples, *params):
        raise NotImplementedError

    def __repr__(self):
        return 'Variable %s of shape %s' % (self.name, str(self.shape))
