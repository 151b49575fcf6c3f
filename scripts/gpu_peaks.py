"""Measured peaks of the box this runs on, the way MEASURED_PEAKS.json was made (torch.matmul 8192^3, best of 10 = burst,
back to back for 4 s = sustained; b.copy_(a) over 1 Gi bf16 elements), plus the TF32 figure BASELINE.md section 4 asks for
before any tensor-roofline fraction is quoted.  Prints one JSON line."""
import json, time
import torch


def mm(dtype, tf32=False, n=8192):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn((n, n), device='cuda', dtype=dtype); b = torch.randn((n, n), device='cuda', dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); reps = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sus = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return best, sus


def copy_bw():
    a = torch.empty((1 << 30,), device='cuda', dtype=torch.bfloat16); b = torch.empty_like(a)
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


if __name__ == '__main__':
    out = {'gpu': torch.cuda.get_device_name(0)}
    out['hbm_copy_gbs'] = copy_bw()
    out['bf16_tflops'], out['bf16_tflops_sustained'] = mm(torch.bfloat16)
    out['tf32_tflops'], out['tf32_tflops_sustained'] = mm(torch.float32, tf32=True)
    out['fp32_simt_tflops'], out['fp32_simt_tflops_sustained'] = mm(torch.float32, tf32=False, n=4096)
    print(json.dumps(out))
