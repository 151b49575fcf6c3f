import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def golden_state(g, t):
    """State vector recorded after t reference steps, as float64 arrays (+ X as int64)."""
    pre = 's%d_' % t
    s = {k[len(pre):]: np.array(g[k], dtype=np.float64) for k in g.files if k.startswith(pre)}
    for k in ('U_hat', 'V_hat', 'log_U_hat', 'log_V_hat'):
        s.pop(k, None)
    s['X'] = g['X'].astype(np.int64)
    return s


def relerr(a, b, floor=1e-6):
    """max |a-b| / (|b| + floor * max|b|): element-wise relative error that does not blow up on entries that
    are (near-)underflow noise next to the array's scale."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if not a.size:
        return 0.0
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor * np.max(np.abs(b)) + 1e-300)))


def relerr_quantile(a, b, q=0.998, floor=1e-6):
    """q-quantile of the element-wise relative error of `relerr`.  For states where a handful of entries are fed by
    terms in float32's DENORMAL range in the reference (exp(lU + lV) ~ 1e-40 keeps a few bits there): those entries of
    the reference carry errors of tens of percent, and only they may differ."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    e = np.abs(a - b) / (np.abs(b) + floor * np.max(np.abs(b)) + 1e-300)
    return float(np.quantile(e, q)), float(e.max())


GOLDEN_CASES = ['zigap_c1', 'gap_c1', 'zigap_ragged', 'gap_ragged', 'zigap_k10',
                'zigap_k32', 'zigap_k64', 'gap_k40']      # the last three: K of configs[3] / configs[4], padded-K plans


@pytest.fixture(scope='session')
def cuda_lib():
    import torch
    from oriana_b200 import _lib
    if not torch.cuda.is_available():
        pytest.skip('no GPU')
    return _lib.load()
