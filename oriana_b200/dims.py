"""Named dimensions and the (samples, distributions, components) view of a node's buffer.

Mirror of oriana/dims.py:11-168.  A relation string such as ``'n,k ~ s,d'`` names the axes of a node's
buffer on the left and tags each of them on the right as sample axis (s), distribution axis (d) or
component axis (c).  The relation knows how to fold a buffer into the canonical 3-D
(n_samples, n_distribs, n_components) array and back; both directions work on numpy arrays and on torch
tensors (device-resident), which is what the nodes of this package hold.
"""
import math

import numpy as np

from .exceptions import IncompatibleShapeException

_TAGS = ('s', 'd', 'c')


def _permute(x, order):
    return x.permute(*order) if hasattr(x, 'permute') else np.transpose(x, order)


class DimRelation:
    """shape <-> (n_samples_per_distrib, n_distribs, n_components) (dims.py:11-48)."""

    def __init__(self, shape, axis_tags):
        self.shape = tuple(int(v) for v in shape)
        self.axis_tags = tuple(axis_tags)
        groups = {t: [i for i, a in enumerate(self.axis_tags) if a == t] for t in _TAGS}
        self._order = groups['s'] + groups['d'] + groups['c']          # canonical axis order
        self._inverse = [self._order.index(i) for i in range(len(self._order))]
        self._grouped_shape = tuple(self.shape[i] for i in self._order)
        self.n_samples_per_distrib = math.prod(self.shape[i] for i in groups['s'])
        self.n_distribs = math.prod(self.shape[i] for i in groups['d'])
        self.n_components = math.prod(self.shape[i] for i in groups['c'])

    @property
    def canonical_shape(self):
        return (self.n_samples_per_distrib, self.n_distribs, self.n_components)

    def reshape_func(self, data):
        """(s, d, c) array -> array of the node's shape (dims.py:136-139)."""
        assert tuple(data.shape) == self.canonical_shape
        return _permute(data.reshape(self._grouped_shape), self._inverse)

    def inv_reshape_func(self, data):
        """array of the node's shape -> (s, d, c) array (dims.py:144-148)."""
        assert tuple(data.shape) == self.shape
        return _permute(data, self._order).reshape(self.canonical_shape)

    def __repr__(self):
        return 'Dimension mapping %s <-> %s' % (str(self.shape), str(self.canonical_shape))


class Dimensions:
    """Mapping from dimension names to sizes; calling it with a relation string builds a DimRelation
    (dims.py:64-151)."""

    def __init__(self, dims):
        self.dims = dict(dims)

    def __call__(self, rel):
        try:
            left, right = rel.split('~')
        except ValueError:
            raise IncompatibleShapeException('Relation "%s" format is not correct.' % rel)
        names = [t.strip() for t in left.strip().split(',')]
        tags = [t.strip() for t in right.strip().split(',')]
        if len(names) != len(tags) or any(t not in _TAGS for t in tags):
            raise IncompatibleShapeException('Relation "%s" format is not correct.' % rel)
        try:
            shape = [self.dims[name] for name in names]
        except KeyError as e:
            raise IncompatibleShapeException('Unknown dimension %s in relation "%s".' % (e, rel))
        return DimRelation(shape, tags)

    def __setitem__(self, key, value):
        self.dims[key] = value   # the reference's setter is a no-op (dims.py:160); this one sets

    def __getitem__(self, key):
        return self.dims[key]
