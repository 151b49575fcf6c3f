"""GPU: `emulate_underflow=True` -- the device models reproduce the float32 exp underflow of the reference's multinomial
step (zigap.py:86-90, gap.py:73-76) -- against the oracle step with the Z sums taken from the sequential float32 loop
(oracle/zloop.c), on the CUDA-core kernels (per term) and on the tcgen05 kernels (per entry)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
PARAMS = ('a1', 'a2', 'b1', 'b2', 'alpha1', 'alpha2', 'beta1', 'beta2')


def _literal_z(lU, lV, X, D_hat=None, quirk=True, dtype=np.float32):
    from oracle import zloop
    if D_hat is None:
        return zloop.gap_z(lU, lV, X.astype(np.float32))
    return zloop.zigap_z(lU, lV, D_hat, X.astype(np.float32), quirk=quirk)


def _state(n, p, K, model, seed):
    """An ordinary initial state except that 40 % of the cells and of the genes have every shape parameter at 0.0143:
    psi(0.0143) = -70.5, so E log U + E log V is about -141 where both meet (0 in the reference's float32 exp) and about
    -70 against an ordinary partner (a normal float32 number)."""
    from oracle import cavi_numpy as cn
    X = cn.synth_counts(n, p, K, seed=seed)
    s = cn.init_state(X, K, np.random.default_rng(seed), model)
    rng = np.random.default_rng(seed + 1)
    low_i = rng.random(n) < 0.4; low_j = rng.random(p) < 0.4
    s['a1'][low_i] = 0.0143; s['b1'][low_j] = 0.0143
    return s, low_i, low_j


@pytest.mark.parametrize('model', ['gap', 'zigap'])
@pytest.mark.parametrize('shape,tensor,tol', [((260, 330, 6), False, 2e-5), ((2048, 1024, 10), True, 3e-3),
                                              ((2048, 1024, 10), 'precise', 1e-4)])
def test_model_step_reproduces_the_reference_underflow(cuda_lib, monkeypatch, model, shape, tensor, tol):
    from oracle import cavi_numpy as cn
    from oriana.models import GaP, ZIGaP
    from oriana.singlecell import CountMatrix
    n, p, K = shape
    s, low_i, low_j = _state(n, p, K, model, seed=n + K)
    cls = ZIGaP if model == 'zigap' else GaP
    kw = dict(k=K, use_factors=False, state=s, compat_quirk=False, tensor=bool(tensor), precise=(tensor == 'precise'))
    m = cls(CountMatrix(s['X']), emulate_underflow=True, **kw)
    plain = cls(CountMatrix(s['X']), **kw)
    assert m.uses_tensor_path == bool(tensor)
    ref = {k: v.copy() for k, v in s.items()}
    monkeypatch.setattr(cn, 'z_expectations', _literal_z)
    for t in range(2):
        alpha1 = ref['alpha1'].copy()
        m.step(); plain.step(); cn.step(ref, quirk=False)
        if t == 0:
            # the reference hands the counts of (low cell, low gene) pairs to nobody: a1 - alpha1 sums to the rest
            got = (m.a1.asarray() - alpha1[None, :]).sum(1)
            want = (ref['a1'] - alpha1[None, :]).sum(1)
            np.testing.assert_allclose(got[low_i], want[low_i], rtol=20 * tol, atol=1e-3)
            lost = s['X'][np.ix_(low_i, low_j)].sum(1)
            assert (lost > 0).any()
            off = (plain.a1.asarray() - alpha1[None, :]).sum(1)
            assert np.all(off[low_i][lost > 0] > want[low_i][lost > 0] + 0.5)      # the exact ratios keep them
        for k in PARAMS + (('pi_d',) if model == 'zigap' else ()):
            e = relerr(getattr(m, k).asarray(), ref[k])
            assert e < tol * (1 if k in ('a1', 'a2', 'b1', 'b2') else 1), (model, t, k, e)


def test_thresholds_do_not_touch_ordinary_states(cuda_lib):
    """With log-expectations of order 1 no entry is anywhere near the rule: the emulating model and the plain one agree
    to the last bit on the CUDA-core kernels."""
    from oracle import cavi_numpy as cn
    from oriana.models import ZIGaP
    from oriana.singlecell import CountMatrix
    X = cn.synth_counts(300, 200, 5, seed=3)
    s = cn.init_state(X, 5, np.random.default_rng(3), 'zigap')
    a = ZIGaP(CountMatrix(X), k=5, use_factors=False, state=s, tensor=False, emulate_underflow=True)
    b = ZIGaP(CountMatrix(X), k=5, use_factors=False, state=s, tensor=False)
    a.step(); b.step()
    for k in PARAMS:
        assert relerr(getattr(a, k).asarray(), getattr(b, k).asarray()) < 1e-6, k


def _literal_sparse_z(lU, lV, S_tilde, S_hat, D_hat, X, dtype=np.float32):
    from oracle import zloop
    return zloop.sparse_z(lU, lV, S_tilde, S_hat, D_hat, X.astype(np.float32))


@pytest.mark.parametrize('shape,tensor,tol', [((260, 330, 6), False, 5e-5), ((2048, 1024, 10), True, 5e-4)])
def test_sparse_model_step_reproduces_the_reference_underflow(cuda_lib, monkeypatch, shape, tensor, tol):
    """SparseZIGaP (sparse_zigap.py:100-116: the same float32 exp, times the mask S_tilde) on the CUDA-core kernels and on
    the tcgen05 kernels (fp32-grade mode, the sparse model's default) against the oracle step with the sequential loop."""
    from oracle import cavi_numpy as cn, sparse_numpy as sn
    from oriana.models import SparseZIGaP
    from oriana.singlecell import CountMatrix
    n, p, K = shape
    X = cn.synth_counts(n, p, K, seed=n + K)
    s = sn.init_state(X, K, np.random.default_rng(n))
    rng = np.random.default_rng(n + 1)
    low_i = rng.random(n) < 0.4; low_j = rng.random(p) < 0.4
    s['a1'][low_i] = 0.0143; s['b1'][low_j] = 0.0143
    kw = dict(k=K, use_factors=False, state=s, tau=0.5, tensor=tensor)
    m = SparseZIGaP(CountMatrix(X), emulate_underflow=True, **kw)
    plain = SparseZIGaP(CountMatrix(X), **kw)
    assert m.uses_tensor_path == tensor
    ref = {k: v.copy() for k, v in s.items()}
    monkeypatch.setattr(sn, 'z_expectations', _literal_sparse_z)
    alpha1 = ref['alpha1'].copy()
    m.step(); plain.step(); sn.step(ref, tau=0.5)
    got = (m.a1.asarray() - alpha1[None, :]).sum(1)
    want = (ref['a1'] - alpha1[None, :]).sum(1)
    np.testing.assert_allclose(got[low_i], want[low_i], rtol=20 * tol, atol=1e-3)
    off = (plain.a1.asarray() - alpha1[None, :]).sum(1)
    lost = X[np.ix_(low_i, low_j)].sum(1)
    assert np.all(off[low_i][lost > 0] > want[low_i][lost > 0] + 0.5)          # the exact ratios keep those counts
    for k in PARAMS + ('pi_d', 'pi_s'):
        e = relerr(getattr(m, k).asarray(), ref[k])
        assert e < tol, (k, e)
    assert relerr(m.p_s.asarray(), ref['p_s']) < 20 * tol
