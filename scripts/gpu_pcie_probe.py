"""Host-to-device copy rates from pinned memory: one large aligned copy, an unaligned start, and the per-slab pattern of the
host-streamed step (bitmap + non-zero bytes + row parameters)."""
import torch, time
dev = torch.device('cuda')
def rate(fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9
N = 1 << 30
h = torch.empty(N + 4096, dtype=torch.uint8, pin_memory=True); h.zero_()
d = torch.empty(N + 4096, dtype=torch.uint8, device=dev)
print('1 GiB aligned        %.1f GB/s' % rate(lambda: d[:N].copy_(h[:N], non_blocking=True), N))
print('1 GiB src +3 bytes   %.1f GB/s' % rate(lambda: d[:N].copy_(h[3:N + 3], non_blocking=True), N))
print('1 GiB src +256 bytes %.1f GB/s' % rate(lambda: d[:N].copy_(h[256:N + 256], non_blocking=True), N))
M = 330 << 20; B = 82 << 20; A = 4 << 20
def slab():
    d[:B].copy_(h[:B], non_blocking=True); d[B:B + M].copy_(h[B:B + M], non_blocking=True)
    d[B + M:B + M + A].copy_(h[B + M:B + M + A], non_blocking=True); d[B + M + A:B + M + 2 * A].copy_(h[B + M + A:B + M + 2 * A], non_blocking=True)
print('slab pattern (82 + 330 + 4 + 4 MiB) %.1f GB/s' % rate(slab, B + M + 2 * A))
d2 = torch.empty(8 << 20, dtype=torch.uint8, device=dev); h2 = torch.empty(8 << 20, dtype=torch.uint8, pin_memory=True)
s2 = torch.cuda.Stream()
def both():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    slab()
print('slab pattern + 8 MiB D2H on another stream %.1f GB/s' % rate(both, B + M + 2 * A))
