"""Step time with and without CUDA-graph replay at the small configurations (launch-bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oriana.models import ZIGaP
from oriana.singlecell import synth_counts_device
for (n, p, K) in [(100, 500, 2), (10_000, 2_000, 10), (100_000, 20_000, 20)]:
    X = synth_counts_device(n, p, K, seed=1)
    out = []
    for graphs in (False, True):
        np.random.seed(0)
        m = ZIGaP(X[:, :p], k=K, use_factors=False, graphs=graphs, trace_cap=600)
        for _ in range(6): m.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 200 if n <= 10_000 else 20
        e0.record()
        for _ in range(steps): m.step()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / steps)
        assert np.all(np.diff(m.elbo_trace) > -1e-6 * np.abs(m.elbo_trace[:-1]))
    print('ZIGaP n=%d p=%d K=%d: %.3f ms/step eager, %.3f ms/step graph replay (%.2fx)' % (n, p, K, out[0], out[1], out[0] / out[1]), flush=True)
