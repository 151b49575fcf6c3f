"""Kernel timings of the two X-streaming passes for ZIGaP / GaP, ELBO on / off (tensor path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oriana.models import GaP, ZIGaP
from oriana.singlecell import synth_counts_device
sizes = [(100_000, 20_000, 20), (250_000, 20_000, 32)]
if 'k64' in sys.argv:
    sizes = [(100_000, 20_000, 48), (100_000, 20_000, 64)]
if 'c5' in sys.argv:       # one rank's share of BASELINE configs[4] over 8 GPUs
    sizes = [(250_000, 30_000, 64)]
quick = 'quick' in sys.argv   # ZIGaP with the ELBO at 250k x 20k, K = 32 only (kernel A/B runs)
if quick:
    sizes = [(250_000, 20_000, 32)]
for (n, p, K) in sizes:
    X = synth_counts_device(n, p, K, seed=1)
    for cls in ((ZIGaP,) if ('c5' in sys.argv or quick) else (ZIGaP, GaP)):
        for elbo in ((True,) if quick else (True, False)):
            np.random.seed(0)
            m = cls(X[:, :p], k=K, use_factors=False, tensor=True, elbo=elbo, precise='precise' in sys.argv,
                    emulate_underflow='underflow' in sys.argv, deterministic='det' in sys.argv)
            for _ in range(2): m.step()
            m.enable_kernel_timing()
            for _ in range(4): m.step()
            kt = m.kernel_times_ms()
            print('%s n=%d p=%d K=%d elbo=%s: rows %.3f ms (%.0f GB/s) genes %.3f ms (%.0f GB/s)' % (
                cls.__name__, n, p, K, elbo, kt['pass_rows'], 4 * n * p / kt['pass_rows'] / 1e6,
                kt['pass_genes'], 4 * n * p / kt['pass_genes'] / 1e6), flush=True)
            del m
    del X
    torch.cuda.empty_cache()
