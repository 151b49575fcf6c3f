"""Which entries send the tensor deviance pass to its float64 redo (config 3, SparseZIGaP after 3 steps)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oriana.models import SparseZIGaP
from oriana.singlecell import synth_counts_device
n, p, K = 100_000, 20_000, 20
X = synth_counts_device(n, p, K, seed=1)
np.random.seed(0)
m = SparseZIGaP(X[:, :p], k=K, use_factors=False)
for _ in range(3): m.step()
m.reconstruction_deviance()
Veff = (m._Vhat[:, :K] * 1.0)
ps = m._ps[:, :K]
b1 = m._b1[:, :K] if hasattr(m, '_b1') else None
print('V_hat zero rows', int((m._Vhat[:, :K].abs().sum(1) == 0).sum()), 'of', p)
print('p_s == 0 fraction', float((ps == 0).float().mean()), ' p_s rows all zero', int((ps.abs().sum(1) == 0).sum()))
U = m._Uhat[m._gen][:, :K]
tot_bad = 0; tot_bad_zero_row = 0; tot_nz = 0
rowzero = (m._Vhat[:, :K].abs().sum(1) == 0)
for r in range(0, n, 8192):
    L = U[r:r + 8192].double() @ m._Vhat[:, :K].double().T
    nz = X[r:r + 8192, :p] != 0
    bad = nz & (L < 1e-30)
    tot_bad += int(bad.sum()); tot_nz += int(nz.sum())
    tot_bad_zero_row += int((bad & rowzero[None, :]).sum())
    if r == 0:
        print('sample: bad with L==0 exactly (fp64):', int((bad & (L == 0)).sum()), 'of', int(bad.sum()))
print('entries', n * p, 'nz', tot_nz, 'bad (nz, rate < 1e-30)', tot_bad, 'of which in all-zero V_hat rows', tot_bad_zero_row)
