"""SparseZIGaP.step() on the device (CUDA-core kernels) at BASELINE configs 2 and 3, the deviance pass, and the numpy
oracle's step at config 2 beside it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oriana.models import SparseZIGaP
from oriana.singlecell import synth_counts_device

for (n, p, K) in [(10_000, 2_000, 10), (100_000, 20_000, 20), (250_000, 20_000, 32)]:
    X = synth_counts_device(n, p, K, seed=1)
    np.random.seed(0)
    m = SparseZIGaP(X[:, :p], k=K, use_factors=False)
    for _ in range(2): m.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps): m.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    m.reconstruction_deviance(); torch.cuda.synchronize()
    t0 = time.perf_counter(); d = m.reconstruction_deviance(); ed = m.explained_deviance(); torch.cuda.synchronize()
    tdev = (time.perf_counter() - t0) * 1e3 / 2
    print('SparseZIGaP n=%d p=%d K=%d: %.2f ms/step = %.3g entries/s; deviance pass %.1f ms (dev=%.6g, explained=%.4f)'
          % (n, p, K, ms, n * p / ms * 1e3, tdev, d, ed), flush=True)
    if n == 10_000:
        from oracle import sparse_numpy as sn
        s = {k: (v.copy() if hasattr(v, 'copy') else v) for k, v in m.state_dict().items()}
        s['X'] = X[:, :p].cpu().numpy().astype(np.int64)
        s['p_d'] = m.D_hat.astype(np.float64)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            sn.step(s)
            t0 = time.perf_counter(); sn.step(s); t1 = time.perf_counter()
        print('  numpy oracle step at the same size: %.2f s = %.3g entries/s (%d host threads)'
              % (t1 - t0, n * p / (t1 - t0), os.cpu_count()), flush=True)
    # the opt-in tensor path: time, and agreement with the CUDA-core path after 3 steps from the same state
    np.random.seed(0)
    ma = SparseZIGaP(X[:, :p], k=K, use_factors=False)
    st = ma.state_dict(); st['X'] = X[:, :p]
    mt = SparseZIGaP(X[:, :p], k=K, use_factors=False, state=st, tensor=True)
    mb = SparseZIGaP(X[:, :p], k=K, use_factors=False, state=st)
    for _ in range(3): mt.step(); mb.step()
    ps_t, ps_s = mt.p_s.asarray(), mb.p_s.asarray()
    rel = lambda a, b: float(np.max(np.abs(a - b) / (np.abs(b) + 1e-6 * np.abs(b).max())))
    print('  tensor vs CUDA-core after 3 steps: a1 %.1e b1(median) %.1e pi_d %.1e | masks differ %.2e, |dp_s|>0.05: %.2e'
          % (rel(mt.a1.asarray(), mb.a1.asarray()), float(np.median(np.abs(mt.b1.asarray() - mb.b1.asarray()) / (np.abs(mb.b1.asarray()) + 1e-12))),
             rel(mt.pi_d.asarray(), mb.pi_d.asarray()), np.mean((ps_t > 0.5) != (ps_s > 0.5)), np.mean(np.abs(ps_t - ps_s) > 0.05)), flush=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps): mt.step()
    e1.record(); torch.cuda.synchronize()
    mst = e0.elapsed_time(e1) / steps
    print('  tensor path: %.2f ms/step = %.3g entries/s (%.1fx)' % (mst, n * p / mst * 1e3, ms / mst), flush=True)
    del m, ma, mt, mb, X
    torch.cuda.empty_cache()
