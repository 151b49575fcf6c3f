"""CPU: the SparseZIGaP restatement (oracle/sparse_numpy.py) against trajectories and convergence metrics recorded
from the UNMODIFIED reference (oracle/make_golden.py, `sparse_*` fixtures).  Groundwork for the next scope row
(the sparsity layer and the deviance metrics the reference's drivers print, main.py:42-44); the
device path does not implement that model yet.

Tolerances: the S-step (sparse_zigap.py:154-163) is a sigmoid of a difference of two large sums, so p_s amplifies the
float32 accumulation-order differences between the numba loop and the BLAS ratio form: 1e-5 after one step, 5e-3 later.
"""
import warnings

import numpy as np
import pytest

from conftest import load_golden, relerr


def _state(g, t):
    pre = 's%d_' % t
    skip = ('deviance', 'explained')       # sparse_gen: the reference's own experiment data (clustering.py:47)
    s = {k[len(pre):]: np.array(g[k], dtype=np.float64) for k in g.files if k.startswith(pre) and k[len(pre):] not in skip}
    s['X'] = g['X'].astype(np.int64)
    return s


@pytest.mark.parametrize('name', ['sparse_k4', 'sparse_ragged'])
def test_sparse_z_kernel_matches_reference_numba_kernel(name):
    from oracle import sparse_numpy as sn
    g = load_golden(name)
    s = _state(g, 0)
    got = sn.z_expectations(g['z_log_U_hat'], g['z_log_Vp_hat'], g['z_S_tilde'], g['z_S_hat'],
                            s['p_d'].astype(np.float32), s['X'])
    for a, key, tol in zip(got, ('z_DSZ', 'z_DZ', 'z_DZl'), (5e-6, 5e-6, 2e-3)):   # third output: two cancelling float32 sums
        assert relerr(a, g[key]) < tol, key


@pytest.mark.parametrize('name', ['sparse_k4', 'sparse_ragged'])
def test_sequential_c_loop_matches_reference_numba_kernel(name):
    """oracle/zloop.c: zl_sparse_z -- same loop order and float32 arithmetic as sparse_zigap.py:100-116 -- against the raw
    kernel call recorded from the reference (the checker of the underflow-emulation tests)."""
    from oracle import zloop
    g = load_golden(name)
    s = _state(g, 0)
    got = zloop.sparse_z(g['z_log_U_hat'], g['z_log_Vp_hat'], g['z_S_tilde'], g['z_S_hat'], s['p_d'].astype(np.float32),
                         s['X'].astype(np.float32))
    for a, key in zip(got[:2], ('z_DSZ', 'z_DZ')):
        assert relerr(a, g[key]) < 2e-6, (key, relerr(a, g[key]))
    # signed terms cancel in sparse_zigap.py:116 (expf of gcc's libm vs numba's differs in the last bit): absolute error
    assert np.max(np.abs(got[2] - g['z_DZl'])) < 3e-6 * np.max(np.abs(g['z_DZl']))


@pytest.mark.parametrize('name', ['sparse_k4', 'sparse_ragged', 'sparse_gen', 'sparse_nmf'])
def test_sparse_trajectory_and_deviance_match_reference(name):
    from oracle import sparse_numpy as sn
    g = load_golden(name)
    s = _state(g, 0)
    steps = [int(t) for t in g['steps']]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for t in range(1, max(steps) + 1):
            sn.step(s, tau=float(g['tau']))
            if t not in steps:
                continue
            want = _state(g, t)
            for k in want:
                if k != 'X':
                    late = 1.5e-2 if name == 'sparse_nmf' else 5e-3      # sparse_nmf = main.py:29 (NMF-seeded: near-dead components)
                    first = 2e-4 if name == 'sparse_nmf' else 1e-5       # beta1 ~ 1e-14 there: psi^-1 of a float32 mean near -1e13
                    assert relerr(s[k], want[k]) < (first if t == 1 else late), (name, t, k)
            dev_ref, expl_ref = float(g['s%d_deviance' % t]), float(g['s%d_explained' % t])
            if abs(dev_ref) < 1e15:            # beyond: a -inf entry was cast to INT64_MIN (quirk Q10), meaningless
                assert abs(sn.reconstruction_deviance(s) - dev_ref) <= 1e-4 * abs(dev_ref), (name, t)
                assert abs(sn.explained_deviance(s) - expl_ref) <= 1e-4 * abs(expl_ref), (name, t)


def test_sparse_fresh_init_runs():
    from oracle import cavi_numpy as cn, sparse_numpy as sn
    X = cn.synth_counts(90, 140, 4, seed=3)
    s = sn.init_state(X, 4, np.random.default_rng(0))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for _ in range(3):
            sn.step(s)
    for k in ('a1', 'a2', 'b1', 'b2', 'p_s', 'pi_s', 'p_d', 'pi_d'):
        assert np.isfinite(s[k]).all(), k
    assert ((s['p_s'] >= 0) & (s['p_s'] <= 1)).all()
