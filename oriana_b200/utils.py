"""Special functions of the CAVI path, evaluated by the CUDA library (oriana/utils.py:9-51).

Same names and argument meaning as the reference; inputs may be numpy arrays, scalars or torch tensors.
numpy in -> numpy out (float64), torch CUDA tensor in -> torch CUDA tensor out.  There is no CPU
implementation here: without the library and a GPU these functions raise.
"""
import numpy as np
import torch

from . import _lib

_OPS = {'digamma': 0, 'digamma_prime': 1, 'inverse_digamma': 2, 'sigmoid': 3, 'logit': 4}


def _special(op, x):
    dev = _lib.require_cuda()
    is_tensor = isinstance(x, torch.Tensor)
    t = x.to(device=dev, dtype=torch.float64) if is_tensor else \
        torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
    t = t.contiguous()
    out = torch.empty_like(t)
    _lib.check(_lib.load().ori_special_f64(_OPS[op], t.data_ptr(), out.data_ptr(), t.numel(), _lib.stream_ptr()))
    if is_tensor:
        return out
    res = out.cpu().numpy()
    return res if res.ndim else float(res)


def logit(x):
    """log(x / (1 - x)) after clipping x to [1e-15, 1 - 1e-15] (utils.py:9-11)."""
    return _special('logit', x)


def sigmoid(x):
    """1 / (1 + exp(-x)) (utils.py:14-15)."""
    return _special('sigmoid', x)


def digamma(x):
    """psi(x) (utils.py:31-32)."""
    return _special('digamma', x)


def digamma_prime(x):
    """psi'(x) (utils.py:35-36)."""
    return _special('digamma_prime', x)


def inverse_digamma(y):
    """Minka's inverse digamma: closed-form start, five Newton steps (utils.py:39-51)."""
    return _special('inverse_digamma', y)


def log(x):
    """log(max(1e-15, x)) (utils.py:18-20); elementwise torch op on the device the input lives on."""
    is_tensor = isinstance(x, torch.Tensor)
    t = x if is_tensor else torch.as_tensor(np.asarray(x, dtype=np.float64))
    out = torch.log(torch.clamp(t, min=1e-15))
    return out if is_tensor else out.numpy()
