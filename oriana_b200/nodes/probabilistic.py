"""Probabilistic nodes (oriana/nodes/probabilistic/*.py).

`Gamma.mean/meanlog` and `Bernoulli.mean` are the expectations the CAVI step consumes
(zigap.py:119-120,127-128,136); they are evaluated by the CUDA library.  `Poisson` and `Multinomial`
are graph decoration in the reference (never evaluated by `step()`); their `mean/sample/logp` are torch ops.
"""
import torch

from .. import _lib
from ..utils import log
from .base import ProbabilisticNode


class Gamma(ProbabilisticNode):
    """Gamma(shape alpha, rate beta) (gamma.py:13-68)."""

    def __init__(self, alpha, beta, rel, **kwargs):
        ProbabilisticNode.__init__(self, alpha, beta, rel=rel, **kwargs)

    def _canon(self, x):
        return x.reshape(1, -1).expand(self.n_samples_per_distrib, -1).unsqueeze(-1)

    def _expect(self, alpha, beta, which):
        dev = _lib.require_cuda()
        a = alpha.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        b = beta.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty_like(a)
        ptrs = [None, None, None]
        ptrs[which] = out.data_ptr()
        _lib.check(_lib.load().ori_gamma_expect_f32(a.data_ptr(), b.data_ptr(), ptrs[0], ptrs[1], ptrs[2],
                                                    a.numel(), _lib.stream_ptr()))
        return out

    def _mean(self, alpha, beta):
        # a / b in float64 (gamma.py:37-46); a single divide, done where the data lives
        return self._canon(alpha.reshape(-1).double() / beta.reshape(-1).double())

    def meanlog(self, recursive=False):
        return self._evaluate(self._meanlog, recursive)

    def _meanlog(self, alpha, beta):
        # psi(float32(a)) - log(float32(b)), float32 result (gamma.py:48-61)
        return self._canon(self._expect(alpha, beta, 1))

    def _sample(self, alpha, beta):
        conc = alpha.reshape(-1).double()
        rate = beta.reshape(-1).double()
        draws = torch.distributions.Gamma(conc, rate).sample((self.n_samples_per_distrib,))
        return draws.unsqueeze(-1)

    def _logp(self, samples, alpha, beta):
        # the reference's own formula AND broadcasting (gamma.py:63-68): `beta` enters as a scale, the normaliser is
        # log(gamma(alpha)) (infinite past alpha = 171.6), and the flat parameter vectors broadcast against the
        # (n, m, 1) samples, so the result is (n, m, m): entry [i, a, b] pairs sample (i, a) with parameters b
        a = alpha.reshape(-1).double(); b = beta.reshape(-1).double()
        s = samples.double()
        out = (a - 1.) * log(s) - (s / b)
        out = out + (-a * log(b) - log(torch.exp(torch.lgamma(a))))
        return out


class Bernoulli(ProbabilisticNode):
    """Bernoulli(pi) (bernoulli.py:12-52)."""

    def __init__(self, pi, rel, **kwargs):
        ProbabilisticNode.__init__(self, pi, rel=rel, **kwargs)

    def _mean(self, pi):
        # float32(pi), tiled over the sample axis (bernoulli.py:41-48)
        p = pi.reshape(1, -1).to(torch.float32)
        return p.expand(self.n_samples_per_distrib, -1).unsqueeze(-1)

    def _sample(self, pi):
        p = pi.reshape(1, -1).double().expand(self.n_samples_per_distrib, -1)
        return torch.bernoulli(p).unsqueeze(-1)

    def _logp(self, samples, pi):
        # bernoulli.py:50-52, with its broadcasting: (n, m, 1) samples against the flat (m,) parameters -> (n, m, m)
        p = pi.reshape(-1).double()
        s = samples.double()
        return s * log(p) + (1. - s) * log(1. - p)


class Poisson(ProbabilisticNode):
    """Poisson(lambda) (poisson.py)."""

    def __init__(self, lambda_, rel, **kwargs):
        ProbabilisticNode.__init__(self, lambda_, rel=rel, **kwargs)

    def _mean(self, lam):
        l = lam.reshape(1, -1).double()
        return l.expand(self.n_samples_per_distrib, -1).unsqueeze(-1)

    def _sample(self, lam):
        l = lam.reshape(1, -1).double().expand(self.n_samples_per_distrib, -1)
        return torch.poisson(l).unsqueeze(-1)

    def _logp(self, samples, lam):
        # poisson.py:64-73: flat result, no log-factorial term, 0 where both are zero, -1000 for a count at rate zero;
        # the result array has the dtype of the samples (`np.empty_like`, :68), so integer counts truncate every entry
        l = lam.reshape(-1).double()
        s = samples.reshape(-1).double()
        zero_rate = torch.where(s > 0, torch.full_like(s, -1000.), torch.zeros_like(s))
        out = torch.where(l > 0, -l + s * log(l), zero_rate)
        return torch.trunc(out) if self.integer_samples else out


class Multinomial(ProbabilisticNode):
    """Multinomial(n, p) over the component axis (multinomial.py); `mean()` = n * p (test/test.py:44-57)."""

    def __init__(self, n, p, rel, **kwargs):
        ProbabilisticNode.__init__(self, n, p, rel=rel, **kwargs)

    def _mean(self, n, p):
        c = self.n_components
        probs = p.reshape(-1, c).double()
        counts = n.reshape(-1, 1).double()
        out = counts * probs
        return out.unsqueeze(0).expand(self.n_samples_per_distrib, -1, -1)

    def _sample(self, n, p):
        c = self.n_components
        probs = p.reshape(-1, c).double()
        counts = n.reshape(-1).long()
        out = torch.zeros(self.n_samples_per_distrib, probs.shape[0], c, dtype=torch.float64, device=probs.device)
        for d in range(probs.shape[0]):
            if counts[d] > 0:
                out[:, d, :] = torch.distributions.Multinomial(int(counts[d]), probs[d]).sample((self.n_samples_per_distrib,))
        return out

    def _logp(self, samples, n, p):
        c = self.n_components
        probs = p.reshape(-1, c).double()
        counts = n.reshape(-1).double()
        lp = torch.lgamma(counts + 1.).unsqueeze(0) - torch.lgamma(samples + 1.).sum(-1) \
            + (samples * log(probs).unsqueeze(0)).sum(-1)
        return lp.unsqueeze(-1).expand(-1, -1, c) / c
