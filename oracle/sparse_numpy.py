"""CPU restatement of the reference's SparseZIGaP iteration and deviance metrics (TEST INFRASTRUCTURE, see
oracle/__init__.py).  SparseZIGaP = ZIGaP + a spike-and-slab layer S on V (`oriana/models/sparse_zigap.py`),
the model the reference's drivers run (`main.py:29`, `experiments/clustering.py:19`).

State vector: X (n,p)  a1,a2 (n,K)  b1,b2 (p,K)  p_d (n,p)  pi_d (p)  p_s (p,K)  pi_s (p)
              alpha1,alpha2,beta1,beta2 (K);  tau is a constructor constant (sparse_zigap.py:17-18).
Pinned against the unmodified reference by the `sparse_*` fixtures of oracle/make_golden.py.
"""
import numpy as np

from . import cavi_numpy as cn


def expectations(s, dtype=np.float32):
    """sparse_zigap.py:198-204."""
    return dict(U_hat=cn.gamma_mean(s['a1'], s['a2']), Vprime_hat=cn.gamma_mean(s['b1'], s['b2']),
                log_U_hat=cn.gamma_meanlog(s['a1'], s['a2'], dtype),
                log_Vprime_hat=cn.gamma_meanlog(s['b1'], s['b2'], dtype),
                D_hat=cn.bernoulli_mean(s['p_d'], dtype), S_hat=cn.bernoulli_mean(s['p_s'], dtype))


def z_expectations(log_U_hat, log_Vp_hat, S_tilde, S_hat, D_hat, X, dtype=np.float32):
    """Ratio-form restatement of the numba triple loop sparse_zigap.py:100-116.

        eU = exp(log U), eV = exp(log V') * S_tilde, den = eU.eV^T (-> 1 where <= 0), R = X * D / den
        DSZ_hat[i,k]           = eU_ik * sum_j R_ij eV_jk S_hat_jk                      :114
        DZ_hat[j,k]            = eV_jk * sum_i R_ij eU_ik                               :115
        DZ_exp_logsum_hat[j,k] = eV_jk * (sum_i R_ij eU_ik logU_ik + logV'_jk sum_i R_ij eU_ik)   :116
    """
    lU = log_U_hat.astype(dtype); lV = log_Vp_hat.astype(dtype)
    eU = cn.centred_exp(lU, dtype)                     # per-row rescaling: every output is a ratio (cavi_numpy.py)
    eV = cn.centred_exp(lV, dtype) * S_tilde.astype(dtype)
    den = eU @ eV.T
    den = np.where(den > 0, den, dtype(1))
    R = X.astype(dtype) * D_hat.astype(dtype) / den
    DSZ = (R @ (eV * S_hat.astype(dtype))) * eU
    RtU = R.T @ eU
    DZ = RtU * eV
    DZl = (R.T @ (eU * lU) + lV * RtU) * eV
    return DSZ.astype(dtype), DZ.astype(dtype), DZl.astype(dtype)


def m_step(s, e):
    """sparse_zigap.py:177-196."""
    s['alpha1'] = cn.clamp(cn.inverse_digamma(np.log(s['alpha2']) + np.mean(e['log_U_hat'], axis=0)))
    s['alpha2'] = cn.clamp(s['alpha1'] / np.mean(e['U_hat'], axis=0))
    s['beta1'] = cn.clamp(cn.inverse_digamma(np.log(s['beta2']) + np.mean(e['log_Vprime_hat'], axis=0)))
    s['beta2'] = cn.clamp(s['beta1'] / np.mean(e['Vprime_hat'], axis=0))
    s['pi_d'] = np.mean(s['p_d'], axis=0)          # :193
    s['pi_s'] = np.mean(s['p_s'], axis=1)          # :196
    return s


def step(s, tau=0.5, dtype=np.float32):
    """One `SparseZIGaP.step()` (base.py:54-56 -> sparse_zigap.py:118-196); `s` is updated in place."""
    e = expectations(s, dtype)
    X = s['X']
    S_tilde = (s['p_s'] > tau).astype(dtype)                                   # :134
    DSZ, DZ, DZl = z_expectations(e['log_U_hat'], e['log_Vprime_hat'], S_tilde, e['S_hat'], e['D_hat'], X, dtype)
    S_hat = e['S_hat']
    V_hat = S_hat * e['Vprime_hat']                                            # :140
    s['a1'] = cn.clamp(s['alpha1'][None, :] + DSZ)                             # :141
    s['a2'] = cn.clamp(s['alpha2'] + e['D_hat'] @ V_hat)                       # :142
    U_hat = cn.gamma_mean(s['a1'], s['a2'])
    DtU = e['D_hat'].T @ U_hat
    s['b1'] = cn.clamp(s['beta1'][None, :] + S_hat * DZ)                       # :149
    s['b2'] = cn.clamp(s['beta2'] + S_hat * DtU)                               # :150
    Vp_hat = cn.gamma_mean(s['b1'], s['b2'])
    tmp = -DZl + np.nan_to_num(DtU * Vp_hat)                                   # :157-158
    pi_s = s['pi_s']
    p_s = np.nan_to_num(cn.sigmoid(cn.logit(pi_s)[:, None] - tmp))             # :159-160
    p_s[pi_s <= 0] = 1e-10                                                     # :161
    p_s[pi_s >= 1] = 1. - 1e-10                                                # :162
    s['p_s'] = p_s
    pi = s['pi_d']
    # :166-167 use the LOCAL V_hat of :140, i.e. the OLD S_hat * OLD Vprime_hat, with the NEW U_hat
    p_d = cn.sigmoid(cn.logit(pi)[None, :] - U_hat @ V_hat.T)
    p_d[:, pi <= 0] = 1e-10
    p_d[:, pi >= 1] = 1. - 1e-10
    p_d[X != 0] = 1. - 1e-10
    s['p_d'] = p_d
    return m_step(s, expectations(s, dtype))


def init_state(X, K, rng, tau=0.5):
    """`use_factors=False` bootstrap (sparse_zigap.py:74-98, base.py:43-52) from our own seeded RNG."""
    n, p = X.shape
    s = dict(X=np.asarray(X))
    s['alpha1'] = rng.gamma(2., size=K); s['alpha2'] = np.ones(K)
    s['beta1'] = rng.gamma(2., size=K); s['beta2'] = np.ones(K)
    s['pi_s'] = rng.random(p); s['pi_d'] = rng.random(p)
    s['a1'] = cn.clamp(rng.gamma(1., size=(n, K))); s['a2'] = np.ones((n, K))
    s['b1'] = cn.clamp(rng.gamma(1., size=(p, K))); s['b2'] = np.ones((p, K))
    s['p_s'] = np.ones((p, K))                                                # :95
    s['p_d'] = (X > 0).astype(np.float64)                                      # :98
    return m_step(s, expectations(s))


# --------------------------------------------------------------------------- deviance metrics
def loglikelihood_X(X, Lambda, pi_d, int_quirk=True):
    """sparse_zigap.py:44-51 (zero-inflated Poisson log-likelihood without the log X! term).

    Reference quirk Q10: the per-entry values are written into `np.empty_like(self.X[:])` (:45), and the X buffer is
    the INTEGER count matrix (`cmatrix.as_array()`, :40), so every entry's log-likelihood is truncated toward zero to
    an int64 before the sum (and a -inf entry becomes INT64_MIN: the metric is meaningless then).  int_quirk=True
    reproduces that; False gives the float64 sum the code presumably meant."""
    Xf = np.asarray(X, dtype=np.float64)
    pi = np.broadcast_to(np.asarray(pi_d, dtype=np.float64)[None, :], Xf.shape)
    z = Xf == 0
    ret = np.empty(Xf.shape, dtype=np.int64 if int_quirk else np.float64)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        ret[z] = np.log(pi[z] * np.exp(-Lambda[z]) + (1. - pi[z]))
        ret[~z] = np.log(pi[~z]) - Lambda[~z] + Xf[~z] * np.log(Lambda[~z])
        return float(ret.sum())


def model_rate(s):
    """UV after base.py:64-67: U_hat (Vprime_hat * S_hat)^T, zero where round(D_hat) == 0."""
    e = expectations(s)
    lam = e['U_hat'] @ (e['Vprime_hat'] * e['S_hat']).T
    lam[np.round(e['D_hat']) == 0] = 0.
    return lam


def reconstruction_deviance(s, int_quirk=True):
    """base.py:58-69."""
    X = s['X'].astype(np.float64)
    return -2. * (loglikelihood_X(X, model_rate(s), s['pi_d'], int_quirk) - loglikelihood_X(X, X, s['pi_d'], int_quirk))


def explained_deviance(s, int_quirk=True):
    """base.py:71-82 (the mask is round(D_hat) == 0, i.e. the D buffer reconstruction_deviance leaves behind; the
    "mean" rate is written into the UV buffer, float64, so it is NOT truncated)."""
    X = s['X'].astype(np.float64)
    ll_sat = loglikelihood_X(X, X, s['pi_d'], int_quirk)
    ll_mean = loglikelihood_X(X, np.broadcast_to(X.mean(axis=0)[None, :], X.shape).copy(), s['pi_d'], int_quirk)
    ll_uv = loglikelihood_X(X, model_rate(s), s['pi_d'], int_quirk)
    return (ll_uv - ll_mean) / (ll_sat - ll_mean)
