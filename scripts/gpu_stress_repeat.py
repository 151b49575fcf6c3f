"""Run the same three CAVI steps many times and report the spread between runs (atomics order only => ~1e-6).
An outlier means a race.  Usage: python scripts/gpu_stress_repeat.py [reps] [n p K]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oriana.models import ZIGaP
from oriana.singlecell import synth_counts_device

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n, p, K = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (4000, 1700, 12)
X = synth_counts_device(n, p, K, seed=8)
np.random.seed(5)
m0 = ZIGaP(X[:, :p], k=K, use_factors=False, tensor=True)
st = m0.state_dict(); st['X'] = X[:, :p]
base = None
worst = {}
for r in range(reps):
    m = ZIGaP(X[:, :p], k=K, use_factors=False, state=st, tensor=True)
    for _ in range(3):
        m.step()
    s = {k: np.asarray(v, dtype=np.float64) for k, v in m.state_dict().items() if k not in ('iterations', 'X')}
    s['elbo'] = np.asarray(m.elbo_trace)
    if base is None:
        base = s; continue
    for k in s:
        d = np.abs(s[k] - base[k]); sc = np.maximum(np.abs(base[k]), 1e-6 * np.abs(base[k]).max())
        e = float((d / sc).max())
        if e > worst.get(k, (0, 0))[0]:
            worst[k] = (e, r)
print('PAIR=%s reps=%d shape=%s' % (os.environ.get('ORI_TC_PAIR', '1'), reps, (n, p, K)))
print('  '.join('%s=%.1e@%d' % (k, v[0], v[1]) for k, v in worst.items()))
