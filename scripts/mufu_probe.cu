// mufu_probe.cu -- issue cost of the MUFU ops of the element-wise stage (ex2 / rcp / lg2 .approx.ftz.f32) on sm_100a:
// cycles per warp instruction and SM sub-partition, with 1, 2 and 4 warps per sub-partition, alone and mixed with FFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mufu_probe scripts/mufu_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP, int NFMA>
__global__ void k_probe(float* out, long long* cyc, int iters, float seed)
{
    float v[16], w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] = seed + 0.001f * (threadIdx.x + 32 * i); w[i] = 0.5f + 0.01f * i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float y;
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
            if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
            if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
            if (OP == 3) {          // the gene pass's mix: ex2 -> fma -> rcp, lg2
                float e, r, l;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[i]));
                const float tz = fmaf(e, w[i], 1.f);
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(tz));
                asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(tz));
                y = r + l;
            }
            if (OP == 4) y = v[i];  // FFMA only
#pragma unroll
            for (int f = 0; f < NFMA; ++f) w[i] = fmaf(w[i], 0.999f, 0.001f * f);
            v[i] = y * 0.5f + 0.25f;
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int NFMA>
void run(const char* name, int mufu_per_elem)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 2000;
    for (int wps = 1; wps <= 4; wps *= 2) {          // warps per SM sub-partition
        const int threads = 128 * wps;
        k_probe<OP, NFMA><<<148, threads>>>(out, cyc, 10, 1.0f);
        k_probe<OP, NFMA><<<148, threads>>>(out, cyc, iters, 1.0f);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double m = 0; for (int i = 0; i < 148; ++i) m += (double)h[i]; m /= 148;
        const double per_elem = m / ((double)iters * 16 * wps);     // cycles per (warp-wide) element step and sub-partition
        printf("%-34s warps/SMSP=%d  %.2f cycles per element step per SMSP", name, wps, per_elem);
        if (mufu_per_elem) printf("  (%.2f per MUFU)", per_elem / mufu_per_elem);
        printf("\n");
    }
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0, 0>("ex2", 1);
    run<1, 0>("rcp", 1);
    run<2, 0>("lg2", 1);
    run<3, 0>("ex2+fma+rcp+lg2", 3);
    run<3, 8>("ex2+fma+rcp+lg2 + 8 ffma", 3);
    run<3, 16>("ex2+fma+rcp+lg2 + 16 ffma", 3);
    run<3, 24>("ex2+fma+rcp+lg2 + 24 ffma", 3);
    run<4, 16>("16 ffma only", 0);
    run<0, 8>("ex2 + 8 ffma", 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
