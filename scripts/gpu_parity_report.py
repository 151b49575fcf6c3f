"""Measured parity of the tensor path (and of the CUDA-core path beside it): the numbers DESIGN.md section 2 quotes.

    python scripts/gpu_parity_report.py [golden] [slabs]

golden: every fixture recorded from the unmodified reference (tests/golden/*.npz, compat_quirk=True), all recorded
        steps up to 50: worst relative error of a1..b2, alpha/beta, pi_d, absolute error of D_hat, and the ELBO of the
        device model against the oracle's float64 ELBO of the reference's own recorded state.
slabs : 2048-row slabs of BASELINE configs 3, 4, 5 (full gene axis, full K) against the oracle port, 6 steps.
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import GOLDEN_CASES, golden_state, load_golden, relerr
from oracle import cavi_numpy as cn
from oriana.models import GaP, ZIGaP
from oriana.singlecell import CountMatrix

FACTORS = ('a1', 'a2', 'b1', 'b2')
HYPER = ('alpha1', 'alpha2', 'beta1', 'beta2')


def make_model(s, quirk, **kw):
    cls = ZIGaP if ('p_d' in s or 'pi_d' in s) else GaP
    return cls(CountMatrix(s['X']), k=s['a1'].shape[1], use_factors=False, state=s, compat_quirk=quirk, **kw)


def golden():
    for name in GOLDEN_CASES:
        g = load_golden(name)
        s = golden_state(g, 0)
        steps = sorted(int(t) for t in g['steps'])
        for tensor in ('precise', True, False):
            if tensor == 'precise' and s['a1'].shape[1] > 32:
                continue
            m = make_model(s, quirk=True, tensor=bool(tensor), precise=(tensor == 'precise'), trace_cap=max(steps) + 8)
            rows = []
            for t in range(1, max(steps) + 1):
                m.step()
                if t in steps:
                    r = golden_state(g, t)
                    ef = max(relerr(getattr(m, k).asarray(), r[k]) for k in FACTORS)
                    eh = max(relerr(getattr(m, k).asarray(), r[k]) for k in HYPER + (('pi_d',) if 'pi_d' in s else ()))
                    ed = float(np.max(np.abs(m.D_hat - r['p_d']))) if 'p_d' in s else 0.0
                    ee = abs(m.elbo() - cn.elbo(r)) / abs(cn.elbo(r))
                    rows.append((t, ef, eh, ed, ee))
            print('%-13s %-7s ' % (name, 'precise' if tensor == 'precise' else ('tensor' if tensor else 'simt')) +
                  ' | '.join('t=%d f %.1e h %.1e D %.1e E %.1e' % r for r in rows), flush=True)


def slabs():
    for cfg, (n, p, K, z) in (('c3', (2048, 20000, 20, 0.5)), ('c4', (2048, 20000, 32, 0.5)), ('c5', (2048, 30000, 64, 0.12))):
        X = cn.synth_counts(n, p, K, seed=0, z=z)
        s = cn.init_state(X, K, np.random.default_rng(0), 'zigap')
        for tensor in ('precise', True, False):
            if tensor == 'precise' and K > 32:
                continue
            m = make_model(s, quirk=False, tensor=bool(tensor), precise=(tensor == 'precise'))
            ref = {k: v.copy() for k, v in s.items()}
            want = [cn.elbo(ref, guard32=True)]
            t0 = time.time()
            out = []
            for t in range(1, 7):
                m.step(); cn.step(ref, quirk=False)
                want.append(cn.elbo(ref))
                ef = max(relerr(getattr(m, k).asarray(), ref[k]) for k in FACTORS)
                eh = max(relerr(getattr(m, k).asarray(), ref[k]) for k in HYPER + ('pi_d',))
                out.append('t=%d f %.1e h %.1e' % (t, ef, eh))
            ed = float(np.max(np.abs(m.D_hat.astype(np.float64) - ref['p_d'])))
            got = m.elbo_trace
            ee = float(np.max(np.abs(got - np.asarray(want)) / np.abs(want)))
            print('%s slab %dx%d K=%d zeros %.2f %-7s %s | D %.1e ELBO trace %.1e (%.0f s)' % (
                cfg, n, p, K, float((X == 0).mean()), 'precise' if tensor == 'precise' else ('tensor' if tensor else 'simt'), ' | '.join(out), ed, ee,
                time.time() - t0),
                flush=True)
            del m


if __name__ == '__main__':
    what = sys.argv[1:] or ['golden', 'slabs']
    if 'golden' in what:
        golden()
    if 'slabs' in what:
        slabs()
