"""Tensor-path diagnostics: parity against the CUDA-core path and the reference fixtures, then timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from conftest import GOLDEN_CASES, golden_state, load_golden, relerr
from oriana.models import GaP, ZIGaP
from oracle import cavi_numpy as cn
P = ('a1', 'a2', 'b1', 'b2', 'alpha1', 'alpha2', 'beta1', 'beta2')
which = sys.argv[1] if len(sys.argv) > 1 else 'all'

if which in ('all', 'parity'):
    for name in GOLDEN_CASES:
        g = load_golden(name); s = golden_state(g, 0)
        cls = ZIGaP if 'p_d' in s else GaP
        K = s['a1'].shape[1]
        for quirk in (True, False):
            mt = cls(s['X'], k=K, use_factors=False, state=s, compat_quirk=quirk, tensor=True)
            ms = cls(s['X'], k=K, use_factors=False, state=s, compat_quirk=quirk, tensor=False)
            assert mt.uses_tensor_path and not ms.uses_tensor_path
            for t in range(1, 6):
                mt.step(); ms.step()
                if t in (1, 5):
                    keys = P + (('pi_d',) if 'pi_d' in s else ())
                    line = ' '.join('%s=%.1e' % (k, relerr(getattr(mt, k).asarray(), getattr(ms, k).asarray())) for k in keys)
                    print(name, 'quirk' if quirk else 'fixed', 't=%d' % t, 'TC-vs-SIMT', line, flush=True)
            et, es = mt.elbo_trace, ms.elbo_trace
            print(name, '  elbo rel diff', np.max(np.abs(et - es) / np.abs(es)), flush=True)
    from oriana.singlecell import synth_counts_device
    for (n, p, K) in ((3000, 1500, 10), (5000, 2100, 32), (1111, 777, 5)):
        X = synth_counts_device(n, p, K, seed=5)
        np.random.seed(3)
        mt = ZIGaP(X[:, :p], k=K, use_factors=False, tensor=True)
        st = mt.state_dict(); st['X'] = X[:, :p]
        ms = ZIGaP(X[:, :p], k=K, use_factors=False, state=st, tensor=False)
        mt = ZIGaP(X[:, :p], k=K, use_factors=False, state=st, tensor=True)
        for t in range(1, 4):
            mt.step(); ms.step()
            keys = P + ('pi_d',)
            print((n, p, K), 't=%d' % t, ' '.join('%s=%.1e' % (k, relerr(getattr(mt, k).asarray(), getattr(ms, k).asarray())) for k in keys), flush=True)
        et, es = mt.elbo_trace, ms.elbo_trace
        print((n, p, K), 'elbo rel diff', np.max(np.abs(et - es) / np.abs(es)), et[-2:], es[-2:], flush=True)

if which in ('all', 'time'):
    from oriana.singlecell import synth_counts_device
    for (n, p, K) in ((10_000, 2_000, 10), (100_000, 20_000, 20), (250_000, 20_000, 32)):
        X = synth_counts_device(n, p, K, seed=1)
        for tensor in ((True,) if 'tconly' in sys.argv else (True, False)):
            np.random.seed(0)
            m = ZIGaP(X[:, :p], k=K, use_factors=False, tensor=tensor)
            for _ in range(2): m.step()
            m.enable_kernel_timing()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(4): m.step()
            ev[1].record(); torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 4
            kt = m.kernel_times_ms()
            print('ZIGaP n=%d p=%d K=%d tensor=%s: %.3f ms/iter  %.3e entries/s  step GB/s %.0f  rows %.2f ms (%.0f GB/s) genes %.2f ms (%.0f GB/s) elbo %s' % (
                n, p, K, tensor, ms, n * p / ms * 1e3, 8 * n * p / ms / 1e6, kt['pass_rows'], 4 * n * p / kt['pass_rows'] / 1e6,
                kt['pass_genes'], 4 * n * p / kt['pass_genes'] / 1e6, m.elbo_trace[-2:]), flush=True)
            del m
        del X
        torch.cuda.empty_cache()
