"""GPU: SparseZIGaP on the device against

  * trajectories and deviance values recorded from the UNMODIFIED reference (`sparse_*` fixtures of
    oracle/make_golden.py: sparse_zigap.py:118-196, base.py:58-82),
  * the oracle restatement (oracle/sparse_numpy.py) on fresh seeded problems, including the float64 deviance.

Tolerances: as in tests/test_oracle_sparse.py -- the S-step is a sigmoid of a difference of two large sums, so p_s
amplifies float32 accumulation-order differences: 2e-5 after one step, 5e-3 later (relative, floor 1e-6 of the
array's scale); the deviance metrics 1e-4 relative (they are sums of integers, see quirk Q10 in oracle/sparse_numpy.py).
"""
import warnings

import numpy as np
import pytest

from conftest import golden_state, load_golden, relerr

pytestmark = pytest.mark.gpu
KEYS = ('a1', 'a2', 'b1', 'b2', 'p_s', 'pi_s', 'pi_d', 'alpha1', 'alpha2', 'beta1', 'beta2')


def _close(got, want, rtol=1e-4):
    """Relative comparison; a log(0) entry makes both sides the same infinity (or the same INT64_MIN-sized number)."""
    if not np.isfinite(want):
        return got == want
    return abs(got - want) <= rtol * abs(want)


def _state(g, t):
    s = golden_state(g, t)
    for k in ('deviance', 'explained'):
        s.pop(k, None)
    return s


def make_model(s, tau, **kw):
    from oriana.models import SparseZIGaP
    from oriana.singlecell import CountMatrix
    return SparseZIGaP(CountMatrix(s['X']), k=s['a1'].shape[1], use_factors=False, state=s, tau=tau, **kw)


@pytest.mark.parametrize('path', ['cuda_core', 'tensor_fp32_grade'])
@pytest.mark.parametrize('name', ['sparse_k4', 'sparse_ragged', 'sparse_gen', 'sparse_nmf'])
def test_sparse_trajectory_and_deviance_match_reference(cuda_lib, name, path):
    """Both kernel families against the reference's recorded trajectories at the SAME tolerances: the CUDA-core kernels
    (what fixtures of this size run by default) and the tcgen05 kernels in their fp32-grade mode (the default from 2^21
    entries on, forced here)."""
    g = load_golden(name)
    m = make_model(_state(g, 0), float(g['tau']), tensor=(path != 'cuda_core'))
    assert m.uses_tensor_path == (path != 'cuda_core') and (path == 'cuda_core' or m.precise)
    if 's0_deviance' in g.files:             # the driver's first print, before any step (clustering.py:20)
        assert abs(m.reconstruction_deviance() - float(g['s0_deviance'])) <= 1e-4 * abs(float(g['s0_deviance']))
    steps = [int(t) for t in g['steps']]
    for t in range(1, max(steps) + 1):
        m.step()
        if t not in steps:
            continue
        want = _state(g, t)
        # sparse_nmf = main.py:29 (NMF-seeded factors with near-dead components): wider late envelope, as for the oracle
        tol = (2e-4 if name == 'sparse_nmf' else 2e-5) if t == 1 else (1.5e-2 if name == 'sparse_nmf' else 5e-3)
        for k in KEYS:
            assert relerr(getattr(m, k).asarray(), want[k]) < tol, (name, t, k)
        assert np.max(np.abs(m.D_hat - want['p_d'])) < (1e-5 if t == 1 else 2e-3), (name, t)
        dev_ref, expl_ref = float(g['s%d_deviance' % t]), float(g['s%d_explained' % t])
        if abs(dev_ref) < 1e15:        # beyond: a -inf entry was cast to INT64_MIN (quirk Q10), meaningless
            assert abs(m.reconstruction_deviance() - dev_ref) <= 1e-4 * abs(dev_ref), (name, t)
            assert abs(m.explained_deviance() - expl_ref) <= 1e-4 * abs(expl_ref), (name, t)


@pytest.mark.parametrize('shape', [(700, 450, 6), (257, 1031, 20), (1500, 90, 32), (300, 210, 40), (129, 77, 64)])
def test_sparse_fresh_problem_matches_oracle(cuda_lib, shape):
    from oracle import cavi_numpy as cn, sparse_numpy as sn
    n, p, K = shape
    X = cn.synth_counts(n, p, K, seed=11)
    s = sn.init_state(X, K, np.random.default_rng(5))
    m = make_model({k: np.array(v, copy=True) for k, v in s.items()}, 0.5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for t in range(1, 4):
            m.step(); sn.step(s, tau=0.5)
            tol = 2e-5 if t == 1 else 5e-3
            for k in KEYS:
                assert relerr(getattr(m, k).asarray(), s[k]) < tol, (shape, t, k)
        # the masks agree entry for entry (no p_s within rounding of tau in these problems)
        assert ((m.p_s.asarray() > 0.5) == (s['p_s'] > 0.5)).all()
        for quirk in (True, False):
            want = sn.reconstruction_deviance(s, int_quirk=quirk)
            assert _close(m.reconstruction_deviance(int_quirk=quirk), want), (shape, quirk)
            want = sn.explained_deviance(s, int_quirk=quirk)
            got = m.explained_deviance(int_quirk=quirk)
            assert _close(got, want) or (np.isnan(want) and np.isnan(got)), (shape, quirk)


def test_sparse_constructor_path_and_snapshot(cuda_lib):
    """`use_factors=False` construction (sparse_zigap.py:74-98, base.py:43-52), a few steps, then a mid-run snapshot
    reloaded into a second model continues identically."""
    from oracle import cavi_numpy as cn
    from oriana.models import SparseZIGaP
    from oriana.singlecell import CountMatrix
    X = cn.synth_counts(300, 200, 4, seed=2)
    np.random.seed(7)
    m = SparseZIGaP(CountMatrix(X), k=4, use_factors=False)
    assert np.all(m.pi_s.asarray() == 1.) and np.all(m.p_s.asarray() == 1.)
    assert np.allclose(m.pi_d.asarray(), (X > 0).mean(axis=0))
    for _ in range(3):
        m.step()
    ps = m.p_s.asarray()
    assert np.isfinite(ps).all() and ((ps >= 0) & (ps <= 1)).all()
    snap = m.state_dict(); snap['X'] = X
    m2 = SparseZIGaP(CountMatrix(X), k=4, use_factors=False, state=snap)
    m.step(); m2.step()
    for k in KEYS:
        assert relerr(getattr(m2, k).asarray(), getattr(m, k).asarray()) < 1e-5, k
    with pytest.raises(ValueError):
        SparseZIGaP(CountMatrix(X), k=65, use_factors=False)
    with pytest.raises(ValueError):
        SparseZIGaP(CountMatrix(X), k=40, use_factors=False, tensor=True)     # tensor plans of this model: k <= 32
    m40 = SparseZIGaP(CountMatrix(X), k=40, use_factors=False)                # 32 < k <= 64: CUDA-core kernels
    assert not m40.uses_tensor_path
    m40.step()
    assert np.isfinite(m40.b1.asarray()).all() and np.isfinite(m40.reconstruction_deviance())


def test_sparse_tensor_paths_track_the_cuda_core_path(cuda_lib):
    """At a size where the tensor kernels are the default: the fp32-grade mode (default) follows the CUDA-core path like
    the CUDA-core path follows the reference (same envelope, masks S_tilde equal almost everywhere); the TF32-operand mode
    (`precise=False`, opt-in) is not comparable step by step (the S update amplifies the TF32 rounding of the gene-side
    sums), so its check is statistical: factors and priors close, masks almost everywhere equal, same deviance."""
    from oriana.models import SparseZIGaP
    from oriana.singlecell import synth_counts_device
    n, p, K = 20_000, 3_000, 8
    X = synth_counts_device(n, p, K, seed=9)
    np.random.seed(3)
    m0 = SparseZIGaP(X[:, :p], k=K, use_factors=False, tensor=False)
    st = m0.state_dict(); st['X'] = X[:, :p]
    ms = SparseZIGaP(X[:, :p], k=K, use_factors=False, state=st, tensor=False)
    mp = SparseZIGaP(X[:, :p], k=K, use_factors=False, state=st)                       # default: tensor, fp32-grade
    mt = SparseZIGaP(X[:, :p], k=K, use_factors=False, state=st, precise=False)        # TF32 operands
    assert mp.uses_tensor_path and mp.precise and mt.uses_tensor_path and not mt.precise and not ms.uses_tensor_path
    for _ in range(3):
        ms.step(); mt.step(); mp.step()
    ps_s, ps_t, ps_p = ms.p_s.asarray(), mt.p_s.asarray(), mp.p_s.asarray()
    # fp32-grade mode.  At this size a handful of (gene, component) pairs sit on the edge of the S update's sigmoid, where ANY
    # change of summation order flips them (measured: |dp_s| > 0.05 on ~1e-5 of the pairs for both tensor modes), and a
    # flipped pair moves the cells it loads on: medians, not maxima
    med = lambda a, b: float(np.median(np.abs(a - b) / (np.abs(b) + 1e-12)))
    for k in ('a1', 'a2', 'b1', 'b2'):
        assert med(getattr(mp, k).asarray(), getattr(ms, k).asarray()) < 1e-5, k
    for k in ('alpha1', 'alpha2', 'beta1', 'beta2'):
        assert relerr(getattr(mp, k).asarray(), getattr(ms, k).asarray()) < 5e-3, k
    assert med(mp.pi_d.asarray(), ms.pi_d.asarray()) < 1e-5
    assert np.quantile(np.abs(mp.pi_s.asarray() - ms.pi_s.asarray()), 0.999) < 1e-3     # pi_s_j = mean_k p_s[j, k]
    assert np.mean((ps_s > 0.5) != (ps_p > 0.5)) < 1e-4
    assert np.mean(np.abs(ps_s - ps_p) > 0.05) < 5e-4
    dsq, dpq = ms.reconstruction_deviance(), mp.reconstruction_deviance()
    assert abs(dpq - dsq) <= 1e-3 * abs(dsq) or not np.isfinite(dsq)
    # TF32-operand mode
    for k in ('a1', 'a2', 'alpha1', 'alpha2', 'beta1', 'beta2', 'pi_d'):
        assert relerr(getattr(mt, k).asarray(), getattr(ms, k).asarray()) < 5e-3, k
    assert np.mean((ps_s > 0.5) != (ps_t > 0.5)) < 5e-3
    assert np.mean(np.abs(ps_s - ps_t) > 0.05) < 2e-2
    same = (ps_s > 0.5) == (ps_t > 0.5)
    b1s, b1t = ms.b1.asarray(), mt.b1.asarray()
    assert np.median(np.abs(b1t - b1s)[same] / (np.abs(b1s)[same] + 1e-12)) < 1e-3
    ds, dt = ms.reconstruction_deviance(int_quirk=False), mt.reconstruction_deviance(int_quirk=False)
    assert abs(dt - ds) <= 5e-3 * abs(ds) or not np.isfinite(ds)


def test_deviance_counts_exactly_zero_rates_like_the_cuda_core_kernel(cuda_lib):
    """Genes whose every component is masked out (S_hat = 0) have an EXACTLY zero rate: an observed count there is log 0
    = -inf, INT64_MIN in the reference's integer buffer (sparse_zigap.py:45).  The tensor deviance pass recognises them
    from the zero patterns of U_hat and V_hat (no float64 redo) and must wrap its first sum exactly like the CUDA-core
    kernel: an odd number of such entries leaves the sign bit set, the rest of the sum agrees."""
    from oriana.models import SparseZIGaP
    from oriana.singlecell import synth_counts_device
    n, p, K = 20_000, 3_000, 8
    X = synth_counts_device(n, p, K, seed=11)
    np.random.seed(5)
    m0 = SparseZIGaP(X[:, :p], k=K, use_factors=False, tensor=False)
    m0.step()
    st = m0.state_dict()
    masked = np.array([3, 64, 65, 1500, 2999])
    st['p_s'][masked] = 0.                               # S_hat = 0 on every component of these genes
    st['X'] = X[:, :p]
    ms = SparseZIGaP(X[:, :p], k=K, use_factors=False, state=st, tensor=False)
    mt = SparseZIGaP(X[:, :p], k=K, use_factors=False, state=st)
    assert mt.uses_tensor_path and not ms.uses_tensor_path
    li_s, li_t = ms._loglikelihood_sums(want_f64=False)[0], mt._loglikelihood_sums(want_f64=False)[0]
    n_inf = int((X[:, :p][:, torch_idx(masked)] != 0).sum())
    assert n_inf > 0
    top = lambda v: (int(v) >> 63) & 1                   # sign bit = parity of the INT64_MIN terms (plus the ordinary sum's sign)
    rest = lambda v: int(v) & ((1 << 63) - 1)
    assert top(li_s[0]) == top(li_t[0])
    assert abs(rest(li_s[0]) - rest(li_t[0])) <= 1e-4 * abs(int(li_s[1]))
    assert abs(int(li_s[1]) - int(li_t[1])) <= 1e-4 * abs(int(li_s[1]))
    assert abs(int(li_s[2]) - int(li_t[2])) <= 1e-4 * abs(int(li_s[2]))


def torch_idx(a):
    import torch
    return torch.as_tensor(a, device='cuda')
