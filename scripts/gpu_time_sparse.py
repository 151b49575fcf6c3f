"""SparseZIGaP.step() on the device at BASELINE configs 2, 3 and a quarter of config 4, per kernel family -- the default
(tcgen05 kernels, fp32-grade mode), the TF32-operand mode, the CUDA-core kernels -- with the agreement of each tensor mode
with the CUDA-core path after 3 steps from the same state, and the deviance pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oriana.models import SparseZIGaP
from oriana.singlecell import synth_counts_device

rel = lambda a, b: float(np.max(np.abs(a - b) / (np.abs(b) + 1e-6 * np.abs(b).max())))
for (n, p, K) in [(10_000, 2_000, 10), (100_000, 20_000, 20), (250_000, 20_000, 32)]:
    X = synth_counts_device(n, p, K, seed=1)
    np.random.seed(0)
    m0 = SparseZIGaP(X[:, :p], k=K, use_factors=False, tensor=False)
    st = m0.state_dict(); st['X'] = X[:, :p]
    del m0
    models = {'cuda-core': dict(tensor=False), 'tensor fp32-grade (default)': dict(), 'tensor tf32 operands': dict(precise=False)}
    ref = None
    for name, kw in models.items():
        m = SparseZIGaP(X[:, :p], k=K, use_factors=False, state=st, **kw)
        for _ in range(3): m.step()
        snap = {k: getattr(m, k).asarray() for k in ('a1', 'b1', 'pi_d', 'p_s')}
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        e0.record()
        for _ in range(steps): m.step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        m.reconstruction_deviance(); torch.cuda.synchronize()
        tdev = 0.0
        for _ in range(3):                      # one sweep of X serves both metrics of a state: time the sweep itself
            m._ll_cache = None
            t0 = time.perf_counter(); d = m.reconstruction_deviance(); ed = m.explained_deviance(); torch.cuda.synchronize()
            tdev += (time.perf_counter() - t0) * 1e3 / 3
        line = 'SparseZIGaP n=%d p=%d K=%d %-28s tensor=%d: %.2f ms/step = %.3g entries/s; deviance pass %.1f ms' % (
            n, p, K, name, m.uses_tensor_path, ms, n * p / ms * 1e3, tdev)
        if ref is None:
            ref = snap
        else:
            line += ' | vs cuda-core after 3 steps: a1 %.1e b1(median) %.1e pi_d %.1e masks differ %.2e |dp_s|>0.05 %.2e' % (
                rel(snap['a1'], ref['a1']), float(np.median(np.abs(snap['b1'] - ref['b1']) / (np.abs(ref['b1']) + 1e-12))),
                rel(snap['pi_d'], ref['pi_d']), np.mean((snap['p_s'] > 0.5) != (ref['p_s'] > 0.5)),
                np.mean(np.abs(snap['p_s'] - ref['p_s']) > 0.05))
        print(line, flush=True)
        del m
    del X
    torch.cuda.empty_cache()
