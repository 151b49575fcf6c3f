"""A few SparseZIGaP steps at config 3 (100k x 20k, K = 20) for an ncu launch list: `tensor` or `simt` as argument."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oriana.models import SparseZIGaP
from oriana.singlecell import synth_counts_device
n, p, K = 100_000, 20_000, 20
X = synth_counts_device(n, p, K, seed=1)
np.random.seed(0)
m = SparseZIGaP(X[:, :p], k=K, use_factors=False, tensor=('tensor' in sys.argv))
for _ in range(3):
    m.step()
d = m.reconstruction_deviance()
torch.cuda.synchronize()
print('tensor path' if m.uses_tensor_path else 'CUDA-core path', 'deviance', d)
