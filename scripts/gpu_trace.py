"""Per-tile timeline of CTA 0 from a -DORI_TC_TRACE build (scripts/build_variant.sh trace "-DORI_TC_TRACE")."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oriana.models import ZIGaP
from oriana.singlecell import synth_counts_device
from oriana_b200 import _lib
n, p, K = 100_000, 20_000, 20
X = synth_counts_device(n, p, K, seed=1)
np.random.seed(0)
m = ZIGaP(X[:, :p], k=K, use_factors=False, tensor=True)
for _ in range(3): m.step()
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = np.zeros((2, 3, 64, 12), dtype=np.int64)
assert lib.ori_debug_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))) == 0
names = ['top', 'xfull', 'sready', 'ld0', 'cmp0', 'st0', 'ld1', 'cmp1', 'st1', 'wait_st', 'arrive']
for ps, pname in enumerate(('rows', 'genes')):
    for w in (0, 1):
        t = buf[ps, w]
        print('== %s EW warp slice %d: per-tile deltas (cycles) between stamps %s' % (pname, w, names))
        for i in range(8, 28):
            d = [int(t[i, k + 1] - t[i, k]) for k in range(10)]
            print('  tile %2d period %5d |' % (i, int(t[i + 1, 0] - t[i, 0])), ' '.join('%5d' % v for v in d))
    t = buf[ps, 2]
    print('== %s MMA warp: top->kfull->S issued->(pready of prev)->tfull->P issued' % pname)
    for i in range(8, 28):
        print('  tile %2d period %5d |' % (i, int(t[i + 1, 0] - t[i, 0])), ' '.join('%5d' % int(t[i, k + 1] - t[i, k]) for k in range(5)))
