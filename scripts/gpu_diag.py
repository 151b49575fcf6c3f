"""Diagnostics for a GPU box: parity error table for every golden case + kernel timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from conftest import GOLDEN_CASES, golden_state, load_golden, relerr
from oriana.models import GaP, ZIGaP
from oracle import cavi_numpy as cn

print(torch.cuda.get_device_name(0), 'cpus', os.cpu_count())
os.system('free -g | head -2; nvidia-smi --query-gpu=memory.total,memory.used --format=csv')
P = ('a1', 'a2', 'b1', 'b2', 'alpha1', 'alpha2', 'beta1', 'beta2')
for name in GOLDEN_CASES:
    g = load_golden(name); s = golden_state(g, 0)
    cls = ZIGaP if 'p_d' in s else GaP
    try:
        m = cls(s['X'], k=s['a1'].shape[1], use_factors=False, state=s, compat_quirk=True)
        steps = [int(t) for t in g['steps']]
        for t in range(1, max(steps) + 1):
            m.step()
            if t in steps:
                r = golden_state(g, t)
                keys = P + (('pi_d',) if 'pi_d' in s else ())
                line = ' '.join('%s=%.1e' % (k, relerr(getattr(m, k).asarray(), r[k])) for k in keys)
                if 'p_d' in s:
                    line += ' D=%.1e' % np.max(np.abs(m.D_hat - r['p_d']))
                print(name, 't=%d' % t, line)
        ref = {k: v.copy() for k, v in s.items()}
        m = cls(s['X'], k=s['a1'].shape[1], use_factors=False, state=s, compat_quirk=False)
        want = [cn.elbo(ref, guard32=True)]
        for t in range(5):
            m.step(); cn.step(ref, quirk=False); want.append(cn.elbo(ref))
        got = m.elbo_trace
        print(name, 'elbo relerr', np.max(np.abs(got - want) / np.abs(want)), 'trace', got[:3], want[:3])
    except Exception as e:
        import traceback; traceback.print_exc()

# timings
from oriana.singlecell import synth_counts_device
for (n, p, K) in ((10_000, 2_000, 10), (100_000, 20_000, 20), (100_000, 20_000, 32)):
    try:
        t0 = time.time(); X = synth_counts_device(n, p, K, seed=1); torch.cuda.synchronize()
        print('synth', n, p, K, '%.2fs' % (time.time() - t0), 'zero frac', float((X[:2000] == 0).float().mean()), 'mean', float(X[:2000].mean()), 'max', float(X[:2000].max()))
        np.random.seed(0)
        m = ZIGaP(X[:, :p], k=K, use_factors=False)
        for _ in range(2): m.step()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5): m.step()
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        print('ZIGaP step n=%d p=%d K=%d: %.3f ms/iter  %.3e entries/s  alg GB/s %.1f' % (n, p, K, ms, n * p / ms * 1e3, 8 * n * p / ms / 1e6))
        print('  elbo trace tail', m.elbo_trace[-3:])
        del m, X
        torch.cuda.empty_cache()
    except Exception as e:
        import traceback; traceback.print_exc()
