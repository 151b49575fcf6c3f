"""Run the same CAVI steps several times from one state and report, per step count, the spread between runs: the only
source of run-to-run variation is the order of the floating-point atomics (the kernels themselves are deterministic), and the
iteration amplifies it step by step.  An outlier against the trend would mean a race.
Usage: python scripts/gpu_stress_repeat.py [reps] [n p K] [simt|precise|det]  (det: ORI_F_DETERMINISTIC, expected spread 0)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oriana.models import ZIGaP
from oriana.singlecell import synth_counts_device

args = [a for a in sys.argv[1:] if a not in ('simt', 'precise', 'det')]
reps = int(args[0]) if len(args) > 0 else 6
n, p, K = (int(a) for a in args[1:4]) if len(args) > 3 else (4000, 1700, 12)
kw = dict(tensor=('simt' not in sys.argv), precise=('precise' in sys.argv), deterministic=('det' in sys.argv))
X = synth_counts_device(n, p, K, seed=8)
np.random.seed(5)
m0 = ZIGaP(X[:, :p], k=K, use_factors=False, **kw)
st = m0.state_dict(); st['X'] = X[:, :p]
del m0
checkpoints = (1, 2, 3, 5, 8)
base = None
worst = {}
for r in range(reps):
    m = ZIGaP(X[:, :p], k=K, use_factors=False, state=st, **kw)
    snaps = {}
    for t in range(1, max(checkpoints) + 1):
        m.step()
        if t in checkpoints:
            snaps[t] = {k: getattr(m, k).asarray() for k in ('a1', 'b1', 'b2', 'alpha1')}
    tr = np.asarray(m.elbo_trace)
    for t in checkpoints:
        snaps[t]['elbo'] = tr[t:t + 1]
    if base is None:
        base = snaps; del m; continue
    for t in checkpoints:
        for k in snaps[t]:
            d = np.abs(snaps[t][k] - base[t][k]); sc = np.maximum(np.abs(base[t][k]), 1e-6 * np.abs(base[t][k]).max())
            e = float((d / sc).max())
            worst[(t, k)] = max(worst.get((t, k), 0.0), e)
    del m
print('%s reps=%d shape=%s' % (kw, reps, (n, p, K)))
for t in checkpoints:
    print('  after %d steps: ' % t + '  '.join('%s=%.1e' % (k, worst[(t, k)]) for k in ('a1', 'b1', 'b2', 'alpha1', 'elbo')))
