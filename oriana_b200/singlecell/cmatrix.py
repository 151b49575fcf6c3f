"""Count-matrix container (oriana/singlecell/cmatrix.py:12-121): the hand-off to the models.

Accepts what the reference accepts (numpy array, pandas DataFrame) plus a torch tensor that may already be
resident in HBM (this rank's row block of a sharded matrix).  Anything else raises DatatypeException
(cmatrix.py:25-29).
"""
import numpy as np
import torch

from ..exceptions import DatatypeException


class CountMatrix:

    def __init__(self, data):
        self._names = None
        if isinstance(data, torch.Tensor):
            if data.dim() != 2:
                raise DatatypeException('Count matrix must be 2-D, got %d-D' % data.dim())
            self._data = data
        elif isinstance(data, np.ndarray):
            if data.ndim != 2:
                raise DatatypeException('Count matrix must be 2-D, got %d-D' % data.ndim)
            self._data = data
        elif hasattr(data, 'values') and hasattr(data, 'columns'):   # pandas DataFrame
            self._names = (list(data.index), list(data.columns))
            self._data = np.asarray(data.values)
        else:
            raise DatatypeException('Incompatible type %s' % type(data))

    def as_array(self):
        """Host numpy array of the counts (cmatrix.py:31-37)."""
        if isinstance(self._data, torch.Tensor):
            return self._data.detach().cpu().numpy()
        return self._data

    def as_tensor(self):
        """The counts as they are stored (device tensor or numpy array), without a copy."""
        return self._data

    @property
    def shape(self):
        return tuple(self._data.shape)

    @property
    def T(self):
        return CountMatrix(self._data.T)

    @staticmethod
    def from_csv(filepath, delimiter=',', has_col_names=True, has_row_names=True):
        import pandas as pd
        df = pd.read_csv(filepath, delimiter=delimiter, header=0 if has_col_names else None,
                         index_col=0 if has_row_names else None)
        return CountMatrix(df)

    def __repr__(self):
        return 'CountMatrix(shape=%s)' % (self.shape,)
