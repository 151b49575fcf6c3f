"""Timeline of the host-streamed step: per slab, when its H2D copies start / end and when its kernels and D2H end, on the
two streams (CUDA events against a common origin).  262144 x 20000, K = 32 (8 slabs), SparseCounts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oriana.singlecell import synth_counts_device
from oriana_b200.host_step import HostStreamedCAVI, SparseCounts, CompactCounts
sys.path.insert(0, ROOT)
import bench
n, p, K = 262144, 20000, 32
X = synth_counts_device(n, p, K, seed=1234)
st = bench.initial_state(n, p, K, 0, n)
mode = sys.argv[1] if len(sys.argv) > 1 else 'sparse'
Xh = SparseCounts.from_tensor(X[:, :p]) if mode == 'sparse' else CompactCounts.from_tensor(X[:, :p])
del X; torch.cuda.empty_cache()

class Timed(HostStreamedCAVI):
    marks = []
    def _copy_slab(self, s, r0, rows):
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        info = HostStreamedCAVI._copy_slab(self, s, r0, rows)
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        self.marks.append([r0, e0, e1])
        return info
    def _finish_upload(self, s, r0, rows, info):
        e2 = torch.cuda.Event(enable_timing=True); e2.record()
        HostStreamedCAVI._finish_upload(self, s, r0, rows, info)
        self.marks[-1].append(e2)
    def _slab_loop(self, body):
        def body2(s, P, r0, rows, stq):
            body(s, P, r0, rows, stq)
            e3 = torch.cuda.Event(enable_timing=True); e3.record()
            self.marks[-1].append(e3)
        HostStreamedCAVI._slab_loop(self, body2)

h = Timed(Xh, K, st, dropout=True)
h.step(); h.step()
import time
Timed.marks = []
origin = torch.cuda.Event(enable_timing=True); origin.record()
t0 = time.perf_counter(); h.step(); dt = time.perf_counter() - t0
torch.cuda.synchronize()
print('%s: step %.1f ms, slab %d rows, h2d %.0f MB per slab' % (mode, dt * 1e3, h.slab, (h.h2d_bytes / 3) / (n / h.slab) / 1e6))
for r0, e0, e1, e2, e3 in Timed.marks:
    print('slab @%7d: H2D %7.2f .. %7.2f   kernels + D2H %7.2f .. %7.2f ms' % (
        r0, origin.elapsed_time(e0), origin.elapsed_time(e1), origin.elapsed_time(e2), origin.elapsed_time(e3)))
