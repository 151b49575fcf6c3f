"""CPU restatement of the reference's CAVI iteration (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to the reference
checkout, AntoinePassemiers/Oriana).  Two flavours are provided:

* ``step(state, ...)``       -- follows the reference's dtypes (float64 parameters, float32
                                log-expectations / D_hat / Z sums) so that it can be pinned
                                against the golden trajectories made by the real reference.
* ``step(..., dtype=np.float64)`` -- the same update order in pure float64 (used as the
                                "true" answer when judging fp32 device results).

The state vector is a plain dict of numpy arrays:
    X (n,p)  a1,a2 (n,K)  b1,b2 (p,K)  p_d (n,p)  pi_d (p)  alpha1,alpha2,beta1,beta2 (K)
GaP has no p_d / pi_d.
"""
import numpy as np
import scipy.special as sp

EPS = 1e-15


# --------------------------------------------------------------------------- special functions
def logit(x):
    """utils.py:9-11 (clip to [1e-15, 1-1e-15] then log-odds)."""
    x = np.clip(x, 1e-15, 1. - 1e-15)
    return np.log(x / (1. - x))


def sigmoid(x):
    """utils.py:14-15."""
    with np.errstate(over='ignore'):
        return 1. / (1. + np.exp(-x))


def digamma(x):
    """utils.py:31-32 (scipy.special.digamma)."""
    return sp.digamma(x)


def digamma_prime(x):
    """utils.py:35-36 (scipy.special.polygamma(1, x))."""
    return sp.polygamma(1, x)


def inverse_digamma(y):
    """utils.py:39-51: Minka's initial guess followed by five Newton steps."""
    y = np.asarray(y, dtype=np.float64)
    x = np.where(y >= -2.22, np.exp(y) + .5, -1. / (y - digamma(1)))
    for _ in range(5):
        x = x - (digamma(x) - y) / digamma_prime(x)
    return x


def clamp(x):
    """zigap.py:117-118 etc.: max(1e-15, nan_to_num(x))."""
    return np.maximum(1e-15, np.nan_to_num(x))


# --------------------------------------------------------------------------- node expectations
def gamma_mean(a, b):
    """gamma.py:37-46: E[U] = a / b in float64."""
    return np.asarray(a, dtype=np.float64) / np.asarray(b, dtype=np.float64)


def gamma_meanlog(a, b, dtype=np.float32):
    """gamma.py:48-61: E[log U] = psi(float32(a)) - log(float32(b)); the result is float32."""
    a = np.asarray(a).astype(dtype)
    b = np.asarray(b).astype(dtype)
    return digamma(a) - np.log(b)


def bernoulli_mean(p, dtype=np.float32):
    """bernoulli.py:41-48: E[D] = float32(p)  (1 - 1e-10 becomes exactly 1.0f)."""
    return np.asarray(p).astype(dtype)


def expectations(s, dtype=np.float32):
    """zigap.py:160-165 / gap.py:131-135: everything that is a pure function of the state."""
    e = dict(U_hat=gamma_mean(s['a1'], s['a2']), V_hat=gamma_mean(s['b1'], s['b2']),
             log_U_hat=gamma_meanlog(s['a1'], s['a2'], dtype),
             log_V_hat=gamma_meanlog(s['b1'], s['b2'], dtype))
    if 'p_d' in s:
        e['D_hat'] = bernoulli_mean(s['p_d'], dtype)
    return e


# --------------------------------------------------------------------------- multinomial latent-count step
EXP_CENTER_LOG2, EXP_FLUSH, EXP_DEAD = 58, -104., -1e4


def centred_exp(log_hat, dtype=np.float32):
    """exp(l - max_k l) * 2^58 per row (2^58 = e^40.2; a power of two keeps every row's leading operand exactly
    representable in tf32 on the device), 0 for components more than e^-104 below their row's largest (the reference's
    float32 exp(lU + lV) underflows below -103.3) and for rows whose largest log-expectation is below -1e4.
    Operands stay in [e^-64, e^40.2]: no product, ratio or accumulated sum leaves float32."""
    l = log_hat.astype(dtype)
    m = l.max(axis=1, keepdims=True)
    with np.errstate(over='ignore', invalid='ignore'):
        rel = l - m
        e = np.ldexp(np.exp(rel), EXP_CENTER_LOG2).astype(dtype)
    return np.where((m > EXP_DEAD) & (rel >= EXP_FLUSH), e, dtype(0)).astype(dtype)


def z_expectations(log_U_hat, log_V_hat, X, D_hat=None, quirk=True, dtype=np.float32):
    """Ratio-form restatement of the numba triple loops zigap.py:79-95 / gap.py:67-80.

        eU = exp(log_U_hat), eV = exp(log_V_hat), den = eU.eV^T (-> 1 where <= 0)
        R  = X / den
        Zi = ((R*D) . eV) * eU                                   zigap.py:93 (D_hat[i, j])
        Zj = (R^T . (eU * D[:, :K])) * eV     when quirk         zigap.py:94 (D_hat[i, k] -- sic)
           = ((R*D)^T . eU) * eV              when not quirk     sparse_zigap.py:115 (correct index)
    GaP (D_hat is None): D == 1 (gap.py:78-80).  The never-read third output
    (zigap.py:95) is not computed.  exp(lu+lv) is evaluated as exp(lu)*exp(lv), each factor rescaled per row
    (`centred_exp`): the ratios the reference forms are invariant to that, and the product stays inside float32 where
    the reference's exp of the SUM does (NMF-initialised factors, base.py:38-40: psi(a1) ~ -1/a1 ~ -100).
    """
    eU = centred_exp(log_U_hat, dtype)
    eV = centred_exp(log_V_hat, dtype)
    X = X.astype(dtype)
    den = eU @ eV.T
    den = np.where(den > 0, den, dtype(1))
    R = X / den
    K = eU.shape[1]
    if D_hat is None:
        Zi = (R @ eV) * eU
        Zj = (R.T @ eU) * eV
    else:
        D = D_hat.astype(dtype)
        RD = R * D
        Zi = (RD @ eV) * eU
        if quirk:
            Zj = (R.T @ (eU * D[:, :K])) * eV
        else:
            Zj = (RD.T @ eU) * eV
    return Zi.astype(dtype), Zj.astype(dtype)


# --------------------------------------------------------------------------- one CAVI iteration
def m_step(s, e):
    """zigap.py:143-158 / gap.py:117-129.  alpha1 uses the OLD alpha2 (zigap.py:146-148)."""
    s['alpha1'] = clamp(inverse_digamma(np.log(s['alpha2']) + np.mean(e['log_U_hat'], axis=0)))
    s['alpha2'] = clamp(s['alpha1'] / np.mean(e['U_hat'], axis=0))
    s['beta1'] = clamp(inverse_digamma(np.log(s['beta2']) + np.mean(e['log_V_hat'], axis=0)))
    s['beta2'] = clamp(s['beta1'] / np.mean(e['V_hat'], axis=0))
    if 'p_d' in s:
        s['pi_d'] = np.mean(s['p_d'], axis=0)  # zigap.py:158
    return s


def step(s, quirk=True, dtype=np.float32):
    """One `FactorModel.step()` (base.py:54-56): E-step (zigap.py:97-141 / gap.py:82-115) then
    M-step.  `s` is updated in place and returned.  `dtype` is the precision of the quantities
    the reference keeps in float32 (log-expectations, D_hat, Z sums); parameters stay float64."""
    zig = 'p_d' in s
    e = expectations(s, dtype)
    X = s['X']
    Zi, Zj = z_expectations(e['log_U_hat'], e['log_V_hat'], X, e.get('D_hat'), quirk, dtype)

    # U_q  (zigap.py:115-120 / gap.py:97-102)
    s['a1'] = clamp(s['alpha1'][None, :] + Zi)
    if zig:
        s['a2'] = clamp(s['alpha2'] + e['D_hat'] @ e['V_hat'])
    else:
        s['a2'] = clamp(np.broadcast_to(s['alpha2'] + e['V_hat'].sum(axis=0), s['a1'].shape).copy())
    U_hat = gamma_mean(s['a1'], s['a2'])

    # V_q, with the NEW U_hat (zigap.py:123-128 / gap.py:105-110)
    s['b1'] = clamp(s['beta1'][None, :] + Zj)
    if zig:
        s['b2'] = clamp(s['beta2'] + e['D_hat'].T @ U_hat)
    else:
        s['b2'] = clamp(np.broadcast_to(s['beta2'] + U_hat.sum(axis=0), s['b1'].shape).copy())
    V_hat = gamma_mean(s['b1'], s['b2'])

    # D_q (zigap.py:131-136); the X != 0 override is applied last
    if zig:
        pi = s['pi_d']
        p_d = sigmoid(logit(pi)[None, :] - U_hat @ V_hat.T)
        p_d[:, pi <= 0] = 1e-10
        p_d[:, pi >= 1] = 1. - 1e-10
        p_d[X != 0] = 1. - 1e-10
        s['p_d'] = p_d

    return m_step(s, expectations(s, dtype))


def init_state(X, K, rng, model='zigap'):
    """The `use_factors=False` bootstrap (zigap.py:55-77, base.py:43-52) from our own seeded RNG:
    a1,b1 ~ Gamma(1), a2=b2=1, p_d=(X>0); prior draws as zigap.py:22,27,32 / gap.py:19-25; then
    one M-step (base.py:52)."""
    n, p = X.shape
    s = dict(X=np.asarray(X))
    if model == 'zigap':
        s['alpha1'] = rng.gamma(2., size=K); s['beta1'] = rng.gamma(2., size=K)
        s['pi_d'] = rng.random(p)
    else:
        s['alpha1'] = np.ones(K); s['beta1'] = np.ones(K)
    s['alpha2'] = np.ones(K); s['beta2'] = np.ones(K)
    s['a1'] = clamp(rng.gamma(1., size=(n, K))); s['a2'] = np.ones((n, K))
    s['b1'] = clamp(rng.gamma(1., size=(p, K))); s['b2'] = np.ones((p, K))
    if model == 'zigap':
        s['p_d'] = (X > 0).astype(np.float64)
    return m_step(s, expectations(s))


# --------------------------------------------------------------------------- ELBO (new; parity unpinned)
def elbo(s, guard32=False):
    """Evidence lower bound of the mean-field family of zigap.py:39-53 for the model of
    zigap.py:21-37, in float64.  The multinomial auxiliary Z is at its
    optimum given the current q(U), q(V), which turns the Poisson term into X*log(den).

      sum_{X>0} [X log den - lgamma(X+1)]  -  sum_ij p_ij (U_hat V_hat^T)_ij
      + sum_ij [p log pi_j + (1-p) log(1-pi_j) - p log p - (1-p) log(1-p)]        (ZIGaP only)
      + sum_ik [a1_ log a2_ - lgamma(a1_) + (a1_-1) ElogU - a2_ U_hat                (prior, a_=alpha)
                + a1 - log a2 + lgamma(a1) + (1-a1) psi(a1)]                          (entropy of q)
      + same for V.
    pi and p are clipped to [1e-15, 1-1e-15].  GaP: p == 1 and no Bernoulli block.

    guard32=True evaluates den the way the reference's Z-step does (zigap.py:86-90): float32
    exp(E log U) . exp(E log V)^T with `den = den if den > 0 else 1`.  The two only differ when the
    float32 exp underflows for every k of an entry, which happens on a random Gamma(1) initial state
    (psi(a1) ~ -1/a1) and never after the first update (a1 >= alpha1).
    """
    X = s['X'].astype(np.float64)
    a1, a2, b1, b2 = (s[k].astype(np.float64) for k in ('a1', 'a2', 'b1', 'b2'))
    U_hat, V_hat = a1 / a2, b1 / b2
    lU = sp.digamma(a1) - np.log(a2)
    lV = sp.digamma(b1) - np.log(b2)
    if guard32:
        den = np.exp(lU.astype(np.float32)) @ np.exp(lV.astype(np.float32)).T
        den = np.where(den > 0, den, np.float32(1)).astype(np.float64)
    else:
        den = np.exp(lU) @ np.exp(lV).T
    nz = X > 0
    out = (X[nz] * np.log(den[nz]) - sp.gammaln(X[nz] + 1.)).sum()
    UV = U_hat @ V_hat.T
    if 'p_d' in s:
        pq = np.clip(s['p_d'].astype(np.float64), 1e-15, 1. - 1e-15)
        pi = np.clip(s['pi_d'].astype(np.float64), 1e-15, 1. - 1e-15)[None, :]
        out -= (pq * UV).sum()
        out += (pq * np.log(pi) + (1. - pq) * np.log(1. - pi)
                - pq * np.log(pq) - (1. - pq) * np.log(1. - pq)).sum()
    else:
        out -= UV.sum()
    for (c1, c2, h1, h2, El, Eh) in ((s['alpha1'], s['alpha2'], a1, a2, lU, U_hat),
                                     (s['beta1'], s['beta2'], b1, b2, lV, V_hat)):
        c1 = c1[None, :]; c2 = c2[None, :]
        out += (c1 * np.log(c2) - sp.gammaln(c1) + (c1 - 1.) * El - c2 * Eh).sum()
        out += (h1 - np.log(h2) + sp.gammaln(h1) + (1. - h1) * sp.digamma(h1)).sum()
    return float(out)


# --------------------------------------------------------------------------- synthetic counts
def synth_counts(n, p, K, seed=0, z=0.5, r=2.0, zinb=True):
    """Moderate-count zero-inflated negative-binomial / Poisson counts:
    U* ~ Gamma(2, .5), V* ~ Gamma(2, .5), Lambda = U* V*^T (mean K); Lambda *= g, g ~ Gamma(r, 1/r);
    L ~ Poisson(Lambda); keep-probability pi_j ~ Beta(1, 1/z - 1) (generation.py:80); X = L * D."""
    rng = np.random.default_rng(seed)
    U = rng.gamma(2., .5, size=(n, K))
    V = rng.gamma(2., .5, size=(p, K))
    lam = U @ V.T
    if zinb:
        lam = lam * rng.gamma(r, 1. / r, size=(n, p))
    L = rng.poisson(lam)
    if z >= 1.0:
        return L.astype(np.int64)
    pi = rng.beta(1., 1. / z - 1., size=p)
    D = rng.random((n, p)) < pi[None, :]
    return (L * D).astype(np.int64)
