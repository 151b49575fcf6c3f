// ew_probe.cu -- the element-wise stage of k_tc_pass (gene pass, dropout, ELBO) in isolation: 8 warps per SM (2 per
// sub-partition, 168-register budget), inputs from shared memory (den / uv as 128-bit loads standing in for tcgen05.ld,
// X as the gene pass's scalar loads), results back to shared memory (standing in for tcgen05.st).  No MMA, no TMA, no
// barriers: what is left is the instruction mix of the per-entry math, so formulations can be compared in seconds.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ew_probe scripts/ew_probe.cu
// Prints cycles per entry row (32 lanes x 1 column) and SM sub-partition for every variant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t tf32_bias(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ float sel_nz_a(float x, float a, float b) {
    float d;
    asm("{\n\t.reg .pred q;\n\tsetp.neu.f32 q, %1, 0f00000000;\n\tselp.f32 %0, %2, %3, q;\n\t}" : "=f"(d) : "f"(x), "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float sel_nz_b(float x, float a, float b) {
    float d;
    asm("{\n\t.reg .pred q;\n\t.reg .f32 t;\n\tabs.f32 t, %1;\n\tsetp.gtu.f32 q, t, 0f00000000;\n\tselp.f32 %0, %2, %3, q;\n\t}"
        : "=f"(d) : "f"(x), "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float is_zero_f(float x) { float d; asm("set.eq.f32.f32 %0, %1, 0f00000000;" : "=f"(d) : "f"(x)); return d; }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// shared layout: den[2 groups][4 quads][256 threads] float4 | uv (same) | x[32 cols][256 threads] float | out R | out D
constexpr int NT = 256;
constexpr uint32_t OFF_DEN = 0, OFF_UV = OFF_DEN + 2 * 4 * NT * 16, OFF_X = OFF_UV + 2 * 4 * NT * 16,
                   OFF_R = OFF_X + 32 * NT * 4, OFF_D = OFF_R + 2 * 4 * NT * 16, SMEM = OFF_D + 2 * 4 * NT * 16;

// V: 0 = the product's gene-pass math (ELBO on)   1 = ELBO off   2 = ELBO by the closed form (no (1-D) e2 term, group max
// instead of the per-entry clamp)   3 = V2 + no tf32 bias adds (compensated truncation folded into the operands)
// 4 = V2 with x as one 128-bit load per 4 entries (row-pass style)  5 = V3 + x as 128-bit loads
// 6 = V0 without the lg2 (knock-out: how much the third MUFU costs here)
template <int V>
__global__ void __launch_bounds__(NT, 1) k_ew(float* out, long long* cyc, int tiles, float lp2j, float cj)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    const int tid = threadIdx.x;
    // plausible contents: den in [0.5, 50], uv*log2e in [-5, 20], 60 % zeros
    for (int i = tid; i < 2 * 4 * NT * 4; i += NT) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        ((float*)(smem + OFF_DEN))[i] = 0.5f + 49.5f * (float)(h & 0xffff) / 65536.f;
        ((float*)(smem + OFF_UV))[i] = -5.f + 25.f * (float)(h >> 16) / 65536.f;
    }
    for (int i = tid; i < 32 * NT; i += NT) {
        uint32_t h = (uint32_t)i * 2246822519u + blockIdx.x * 7919u;
        h ^= h >> 15; h *= 2654435761u; h ^= h >> 13;
        ((float*)(smem + OFF_X))[i] = (h & 0xff) < 154 ? 0.f : (float)(1 + ((h >> 8) & 7));
    }
    __syncthreads();
    const float ulim = fminf(127.f, 127.f + lp2j);
    const float lp2f = fminf(lp2j, 3.0e38f);
    float cs = 0.f, xl = 0.f, ent = 0.f, dmin_all = 1.f, umax_all = 0.f;
    constexpr bool X128 = (V == 4 || V == 5);
    constexpr bool ELBO = (V != 1);
    constexpr bool CLOSED = (V == 2 || V == 3 || V == 4 || V == 5);
    constexpr bool NOBIAS = (V == 3 || V == 5);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            uint32_t dr[16], ur[16];
            float x[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 a = lds128(sb + OFF_DEN + ((g * 4 + q) * NT + tid) * 16);
                const float4 b = lds128(sb + OFF_UV + ((g * 4 + q) * NT + tid) * 16);
                dr[4 * q] = __float_as_uint(a.x); dr[4 * q + 1] = __float_as_uint(a.y); dr[4 * q + 2] = __float_as_uint(a.z); dr[4 * q + 3] = __float_as_uint(a.w);
                ur[4 * q] = __float_as_uint(b.x); ur[4 * q + 1] = __float_as_uint(b.y); ur[4 * q + 2] = __float_as_uint(b.z); ur[4 * q + 3] = __float_as_uint(b.w);
            }
            if (X128) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 a = lds128(sb + OFF_X + ((g * 4 + q) * NT + tid) * 16);
                    x[4 * q] = a.x; x[4 * q + 1] = a.y; x[4 * q + 2] = a.z; x[4 * q + 3] = a.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) x[e] = lds32(sb + OFF_X + ((g * 16 + e) * NT + tid) * 4);
            }
            float dmin = 1.f, umax = 0.f, g_cs = 0.f, g_xl = 0.f, g_ent = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float den = __uint_as_float(dr[e]);
                const float xe = x[e];
                dmin = fminf(dmin, den);
                float uvp = __uint_as_float(ur[e]);
                if (ELBO && !CLOSED) uvp = fminf(uvp, ulim);
                if (CLOSED) umax = fmaxf(umax, uvp);
                const float tz = fmaf(ex2_approx(uvp), cj, 1.f);
                const float tt = sel_nz_a(xe, den, tz);
                const float r = rcp_approx(tt);
                dr[e] = NOBIAS ? __float_as_uint(xe * r) : tf32_bias(xe * r);
                const float D = sel_nz_b(xe, 1.f, r);
                ur[e] = NOBIAS ? __float_as_uint(D) : tf32_bias(D);
                g_cs += D;
                if (ELBO) {
                    const float l2 = (V == 6) ? tt : lg2_approx(tt);
                    g_xl = fmaf(xe, l2, g_xl);
                    g_ent = fmaf(is_zero_f(xe), l2, g_ent);
                    if (!CLOSED) {
                        const float e2 = uvp - lp2f;
                        const float w = 1.f - D;
                        g_ent = fmaf(-w, e2, g_ent);
                    }
                }
            }
            cs += g_cs; xl += g_xl; ent += g_ent;
            dmin_all = fminf(dmin_all, dmin); umax_all = fmaxf(umax_all, umax);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                sts128(sb + OFF_R + ((g * 4 + q) * NT + tid) * 16, dr[4 * q], dr[4 * q + 1], dr[4 * q + 2], dr[4 * q + 3]);
                sts128(sb + OFF_D + ((g * 4 + q) * NT + tid) * 16, ur[4 * q], ur[4 * q + 1], ur[4 * q + 2], ur[4 * q + 3]);
            }
        }
        __syncwarp();
    }
    const long long t1 = clock64();
    out[blockIdx.x * NT + tid] = cs + xl + ent + dmin_all + umax_all;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * NT * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaFuncSetAttribute(k_ew<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    const int tiles = 4000;
    k_ew<V><<<148, NT, SMEM>>>(out, cyc, 10, 1.5f, 0.35f);
    k_ew<V><<<148, NT, SMEM>>>(out, cyc, tiles, 1.5f, 0.35f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += (double)h[i]; m /= 148;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_ew<V>);
    printf("%-70s %6.2f cycles per entry row per SMSP  (%d regs)\n", name, m / ((double)tiles * 64), fa.numRegs);
    cudaFree(out); cudaFree(cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
    run<0>("V0 product gene-pass math, ELBO on");
    run<1>("V1 ELBO off");
    run<6>("V6 V0 without MUFU.LG2");
    run<2>("V2 ELBO closed form (no (1-D) e2 term, group max for the clamp)");
    run<3>("V3 V2 + no tf32 bias adds");
    run<4>("V4 V2 + X as 128-bit loads");
    run<5>("V5 V3 + X as 128-bit loads");
    return 0;
}
