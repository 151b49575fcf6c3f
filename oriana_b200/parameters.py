"""`Parameter` -- the reference's array container (oriana/parameters.py:8-32), backed by a torch tensor
that lives in HBM when a GPU is present.

Semantics kept: float64 storage by default (parameters.py:11), `p[key]` reads, `p[key] = v` writes,
`.shape`, `.asarray()`, `.buffer`.  Reads return HOST numpy arrays (copies): `p[:] += x` still works because
Python re-assigns through `__setitem__`.  `.tensor` is the zero-copy device view.
"""
import numpy as np
import torch


def default_device():
    return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')


def as_tensor(data, dtype=torch.float64, device=None):
    device = device or default_device()
    if isinstance(data, Parameter):
        data = data.tensor
    if isinstance(data, torch.Tensor):
        return data.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(data, dtype=np.float64), dtype=dtype, device=device)


class Parameter:

    def __init__(self, data, device=None, dtype=torch.float64):
        self._t = as_tensor(data, dtype=dtype, device=device).clone() if isinstance(data, torch.Tensor) \
            else as_tensor(data, dtype=dtype, device=device)

    # -- reference API ---------------------------------------------------------------------------
    def asarray(self):
        return self._t.detach().cpu().numpy()

    def __getitem__(self, key):
        return self.asarray()[key]

    def __setitem__(self, key, value):
        if isinstance(value, Parameter):
            value = value.tensor
        if not isinstance(value, torch.Tensor):
            value = torch.as_tensor(np.asarray(value, dtype=np.float64))
        self._t[key] = value.to(device=self._t.device, dtype=self._t.dtype)

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def buffer(self):
        return self.asarray()

    @buffer.setter
    def buffer(self, data):
        self._t = as_tensor(data, dtype=self._t.dtype, device=self._t.device)

    # -- device view --------------------------------------------------------------------------------
    @property
    def tensor(self):
        return self._t

    def __array__(self, dtype=None, copy=None):
        a = self.asarray()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self._t.shape[0]

    def __repr__(self):
        return 'Parameter(shape=%s, device=%s)' % (self.shape, self._t.device)
