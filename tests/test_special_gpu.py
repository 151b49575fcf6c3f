"""GPU: device special functions and node expectations against the reference's own known answers
(test/test.py:13-32, :60-79) and the values its utils produced (tests/golden/special.npz)."""
import numpy as np
import pytest
from numpy.testing import assert_almost_equal

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def test_sigmoid_logit_roundtrip(cuda_lib):
    from oriana.utils import logit, sigmoid
    x = np.asarray([-2.3, 1.5, 0.45, -0.78, 5.3, -.2, 0.])
    assert_almost_equal(logit(sigmoid(x)), x)                       # test/test.py:13-15
    x = np.asarray([0.45, 0.001, 0.9987, 0.63, 0.745, 0.521, 0.32])
    assert_almost_equal(sigmoid(logit(x)), x)                       # test/test.py:18-20


def test_digamma_roundtrip(cuda_lib):
    from oriana.utils import digamma, inverse_digamma
    x = np.asarray([0.54, 6.2, 1.2, 0.3, 7.9, 4.5, 2.1])
    assert_almost_equal(x, inverse_digamma(digamma(x)))             # test/test.py:23-26
    assert_almost_equal(x, digamma(inverse_digamma(x)))             # test/test.py:29-32


def test_special_values_match_scipy_backed_reference(cuda_lib):
    from oriana.utils import digamma, digamma_prime, inverse_digamma, logit, sigmoid
    g = load_golden('special')
    assert np.max(np.abs(digamma(g['x']) - g['digamma']) / (1 + np.abs(g['digamma']))) < 1e-13
    assert relerr(digamma_prime(g['x']), g['trigamma'], floor=0) < 1e-12
    assert relerr(inverse_digamma(g['y']), g['inverse_digamma'], floor=0) < 1e-10
    assert np.max(np.abs(sigmoid(g['z']) - g['sigmoid'])) < 1e-15
    assert np.max(np.abs(logit(g['q']) - g['logit'])) < 1e-12


def test_gamma_mean_and_meanlog(cuda_lib):
    """test/test.py:60-79."""
    from oriana import Dimensions, Parameter
    from oriana.nodes import Gamma
    from oriana.utils import digamma
    alpha1 = Parameter([[2.1, 1.8], [0.7, 2.3]])
    alpha2 = Parameter(np.ones((2, 2)))
    dims = Dimensions({'n': 2, 'm': 2, 'k': 2})
    gamma = Gamma(alpha1, alpha2, dims('n,m,k ~ d,s,d'))
    y = np.asarray([[[2.1, 1.8], [2.1, 1.8]], [[0.7, 2.3], [0.7, 2.3]]])
    assert_almost_equal(gamma.mean(), y)
    x = gamma.meanlog()
    assert x.dtype == np.float32                                     # gamma.py:56-57
    assert_almost_equal(x, digamma(y), decimal=6)


def test_gamma_expect_kernel_against_oracle(cuda_lib):
    """gamma.py:37-61 over the whole float32 range the clamps allow (zigap.py:117-118)."""
    import torch
    from oriana_b200 import _lib
    from oracle import cavi_numpy as cn
    rng = np.random.default_rng(0)
    a1 = np.concatenate([10 ** rng.uniform(-15, 6, 4000), [1e-15, 1.0, 1.4616321, 3e38]]).astype(np.float32)
    a2 = np.concatenate([10 ** rng.uniform(-15, 6, 4000), [1e-15, 1.0, 2.0, 1.0]]).astype(np.float32)
    ta, tb = torch.as_tensor(a1).cuda(), torch.as_tensor(a2).cuda()
    E, El, eE = (torch.empty_like(ta) for _ in range(3))
    _lib.check(cuda_lib.ori_gamma_expect_f32(ta.data_ptr(), tb.data_ptr(), E.data_ptr(), El.data_ptr(),
                                             eE.data_ptr(), ta.numel(), _lib.stream_ptr()))
    ref_E = cn.gamma_mean(a1, a2); ref_l = cn.gamma_meanlog(a1, a2)
    assert relerr(E.cpu().numpy(), ref_E.astype(np.float32), floor=0) < 2e-7
    assert np.max(np.abs(El.cpu().numpy() - ref_l) / (1 + np.abs(ref_l))) < 3e-7
    fin = np.abs(ref_l) < 80
    assert relerr(eE.cpu().numpy()[fin], np.exp(ref_l[fin].astype(np.float64)), floor=0) < 2e-5
