import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cavi_numpy as cn
from oriana.models import ZIGaP
from oriana.singlecell import CountMatrix
X = cn.synth_counts(400, 300, 5, seed=9)
np.random.seed(0)
m = ZIGaP(CountMatrix(X), k=5, use_factors=True, nmf='host')
print('init: zeros in nmf U', (m._nmf_U == 0).sum(), 'V', (m._nmf_V == 0).sum(), 'zero rows U', (m._nmf_U.sum(1) == 0).sum(), 'zero rows V', (m._nmf_V.sum(1) == 0).sum())
print('elbo0', m.elbo())
p, KP = m.p, m._KP
def dump(tag):
    torch.cuda.synchronize()
    print(tag, 'parts', m._red64[p + 2 * KP:].cpu().numpy())
    print(tag, 'SlogU', m._red64[p:p + KP].cpu().numpy()[:5], 'SU', m._red64[p + KP:p + 2 * KP].cpu().numpy()[:5])
    print(tag, 'gsum', m._gsum.cpu().numpy())
    print(tag, 'scal', m._scal.cpu().numpy()[:8])
    print(tag, 'hyper', m._hyper.cpu().numpy())
    for k in ('a1', 'a2', 'b1', 'b2'):
        v = getattr(m, k).asarray(); print(tag, k, np.isfinite(v).all(), v.min(), v.max())
dump('t0')
m.step()
dump('t1-before-finalize')
print('elbo1', m.elbo())
dump('t1')
