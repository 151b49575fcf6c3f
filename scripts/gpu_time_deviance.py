"""CUDA-event time of the deviance sweep (ori_deviance_sums, integer mode) on the tensor path and on the CUDA-core kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oriana.models import SparseZIGaP
from oriana.singlecell import synth_counts_device
for (n, p, K) in [(100_000, 20_000, 20), (250_000, 20_000, 32)]:
    X = synth_counts_device(n, p, K, seed=1)
    np.random.seed(0)
    for kw in (dict(), dict(tensor=False)):
        m = SparseZIGaP(X[:, :p], k=K, use_factors=False, **kw)
        for _ in range(3): m.step()
        m.reconstruction_deviance()
        pi = m._current_pi()
        out_i = torch.zeros((3,), dtype=torch.int64, device='cuda')
        ts = []
        for _ in range(4):
            out_i.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m._call('ori_deviance_sums', m._gen, pi.data_ptr(), m._col_mean.data_ptr(), out_i.data_ptr(), None)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('n=%d p=%d K=%d tensor=%d: deviance sweep %s ms, sums %s' % (n, p, K, m.uses_tensor_path, ['%.2f' % t for t in ts], out_i.cpu().numpy()), flush=True)
        del m
    del X; torch.cuda.empty_cache()
