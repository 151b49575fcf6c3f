// tc_probe.cu -- building-block checks for the tensor-core path, run on a B200 before the real kernels:
//   T1  TMA(SW128) -> tcgen05.mma kind::tf32, A and B K-major from smem           D1[128x64] = A1[128x32] . B1[64x32]^T
//   T2  multi-chunk K-major A and B (contraction spans two 128-byte swizzle chunks)   D2[128x32] = A2[128x64] . B2T[32x64]^T
//   T3  A from TMEM (tcgen05.st -> mma TS) with the B of T2                          D3 == D2
// (MN-major tf32 operands need the SWIZZLE_128B_ATOM_32B layout, which no K-major operand accepts, so the
//  kernels keep transposed copies of the factor arrays and use K-major descriptors everywhere.)
//   T4  what the tensor core does with fp32 inputs that are not tf32-representable (truncate or round)
// Inputs of T1-T3 are tf32-exact, so the results must match the CPU bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_probe scripts/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../oriana_b200/csrc/tc_ptx.cuh"

using namespace tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Maps { CUtensorMap a1, b1, a2, b2; };

// mode 1: T1, mode 2: T2, mode 3: T3 (A2 via TMEM), mode 4: T4 (A1 = const nonrepresentable)
__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ Maps maps, int mode, const float* __restrict__ A2g,
                                               float* __restrict__ out, int out_cols)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* sA = (float*)smem;                    // up to 2 chunks [128 x 32] = 32 KB
    float* sB = (float*)(smem + 32768);          // [64 x 32] = 8 KB, or 2 chunks [32 x 32] of B2T
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        if (mode == 1 || mode == 4) {
            mbar_expect_tx(&bar_load, 16384 + 8192);
            tma_load_2d(sA, &maps.a1, &bar_load, 0, 0);
            tma_load_2d(sB, &maps.b1, &bar_load, 0, 0);
        } else {
            mbar_expect_tx(&bar_load, (mode == 2 ? 32768 : 0) + 8192);
            if (mode == 2) {
                tma_load_2d(sA, &maps.a2, &bar_load, 0, 0);
                tma_load_2d(sA + 4096, &maps.a2, &bar_load, 32, 0);
            }
            tma_load_2d(sB, &maps.b2, &bar_load, 0, 0);
            tma_load_2d(sB + 1024, &maps.b2, &bar_load, 32, 0);
        }
    }
    if (mode == 3) {   // every thread writes its row of A2 (64 values) into TMEM columns [64, 128)
        const int row = tid;
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t v[16];
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(A2g[row * 64 + c0 + i]);
            tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 64 + c0, v);
        }
        tmem_wait_st();
        tc_fence_before();
    }
    __syncthreads();

    if (tid == 0) {
        mbar_wait(&bar_load, 0, 1);
        tc_fence_after();
        if (mode == 1 || mode == 4) {
            const uint32_t idesc = make_idesc_tf32(128, 64, false, false);
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = make_smem_desc(smem_u32(sA) + k * 32, 16, 1024);
                const uint64_t bd = make_smem_desc(smem_u32(sB) + k * 32, 16, 1024);
                mma_tf32_ss(tmem, ad, bd, idesc, k > 0);
            }
        } else {
            const uint32_t idesc = make_idesc_tf32(128, 32, false, false);
            for (int ks = 0; ks < 8; ++ks) {
                const uint64_t bd = make_smem_desc(smem_u32(sB) + (ks >> 2) * 4096 + (ks & 3) * 32, 16, 1024);
                if (mode == 2) {
                    const uint64_t ad = make_smem_desc(smem_u32(sA) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
                    mma_tf32_ss(tmem, ad, bd, idesc, ks > 0);
                } else {
                    mma_tf32_ts(tmem, tmem + 64 + ks * 8, bd, idesc, ks > 0);
                }
            }
        }
        tc_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0, 2);
    tc_fence_after();
    for (int c0 = 0; c0 < out_cols; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
        for (int i = 0; i < 16; ++i) out[tid * out_cols + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

static float q(int v) { return 0.25f * (float)v; }

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s sm_%d%d\n", prop.name, prop.major, prop.minor);
    std::vector<float> A1(128 * 32), B1(64 * 32), A2(128 * 64), B2(64 * 32), A4(128 * 32), B4(64 * 32, 1.0f);
    srand(1);
    for (auto& v : A1) v = q(rand() % 17 - 8);
    for (auto& v : B1) v = q(rand() % 17 - 8);
    for (auto& v : A2) v = q(rand() % 17 - 8);
    for (auto& v : B2) v = q(rand() % 17 - 8);
    const float odd = 1.0f + 3.0f / 4096.0f;   // 1 + 0.75 * 2^-10: truncation -> 1, round-to-nearest -> 1 + 2^-10
    for (auto& v : A4) v = odd;
    float *dA1, *dB1, *dA2, *dB2, *dA4, *dB4, *dout;
    CK(cudaMalloc(&dA1, A1.size() * 4)); CK(cudaMalloc(&dB1, B1.size() * 4)); CK(cudaMalloc(&dA2, A2.size() * 4));
    CK(cudaMalloc(&dB2, B2.size() * 4)); CK(cudaMalloc(&dA4, A4.size() * 4)); CK(cudaMalloc(&dB4, B4.size() * 4));
    CK(cudaMalloc(&dout, 128 * 64 * 4));
    CK(cudaMemcpy(dA1, A1.data(), A1.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB1, B1.data(), B1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA2, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB2, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA4, A4.data(), A4.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB4, B4.data(), B4.size() * 4, cudaMemcpyHostToDevice));

    Maps m, m4;
    bool ok = make_tmap_f32(&m.a1, dA1, 128, 32, 32, 32, 128) && make_tmap_f32(&m.b1, dB1, 64, 32, 32, 32, 64) &&
              make_tmap_f32(&m.a2, dA2, 128, 64, 64, 32, 128) && make_tmap_f32(&m.b2, dB2, 32, 64, 64, 32, 32);
    m4 = m;
    ok = ok && make_tmap_f32(&m4.a1, dA4, 128, 32, 32, 32, 128) && make_tmap_f32(&m4.b1, dB4, 64, 32, 32, 32, 64);
    if (!ok) { printf("tensor map creation failed\n"); return 2; }
    const int smem_bytes = 32768 + 8192 + 1024;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));

    int fails = 0;
    std::vector<float> out(128 * 64);
    auto run = [&](int mode, const Maps& mm, int cols) {
        CK(cudaMemset(dout, 0xff, 128 * 64 * 4));
        k_probe<<<1, 128, smem_bytes>>>(mm, mode, dA2, dout, cols);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T%d: kernel failed: %s\n", mode, cudaGetErrorString(e)); exit(3); }
        CK(cudaMemcpy(out.data(), dout, 128 * cols * 4, cudaMemcpyDeviceToHost));
    };
    // T1
    run(1, m, 64);
    {
        int bad = 0; double maxd = 0;
        for (int i = 0; i < 128; ++i) for (int j = 0; j < 64; ++j) {
            float r = 0; for (int k = 0; k < 32; ++k) r += A1[i * 32 + k] * B1[j * 32 + k];
            double d = fabs((double)r - out[i * 64 + j]); if (d > maxd) maxd = d;
            if (d != 0 && bad++ < 5) printf("  T1 mismatch [%d,%d] got %g want %g\n", i, j, out[i * 64 + j], r);
        }
        printf("T1 (SS, K-major A/B, SW128 via TMA): %s  bad=%d maxdiff=%g\n", bad ? "FAIL" : "PASS", bad, maxd); fails += bad != 0;
    }
    std::vector<float> ref2(128 * 32);
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 32; ++j) {
        float r = 0; for (int k = 0; k < 64; ++k) r += A2[i * 64 + k] * B2[j * 64 + k];
        ref2[i * 32 + j] = r;
    }
    for (int mode = 2; mode <= 3; ++mode) {
        run(mode, m, 32);
        int bad = 0; double maxd = 0;
        for (int i = 0; i < 128 * 32; ++i) {
            double d = fabs((double)ref2[i] - out[i]); if (d > maxd) maxd = d;
            if (d != 0 && bad++ < 5) printf("  T%d mismatch [%d,%d] got %g want %g\n", mode, i / 32, i % 32, out[i], ref2[i]);
        }
        printf("T%d (%s A, 2-chunk K-major B): %s  bad=%d maxdiff=%g\n", mode, mode == 2 ? "2-chunk K-major smem" : "TMEM", bad ? "FAIL" : "PASS", bad, maxd);
        fails += bad != 0;
    }
    run(4, m4, 64);
    printf("T4 fp32 operand %.10f x 32: got %.10f  (truncate -> 32, round-to-nearest -> %.10f, exact fp32 -> %.10f)\n",
           odd, out[0], 32.0f * (1.0f + 1.0f / 1024.0f), 32.0f * odd);
    printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
    return fails ? 1 : 0;
}
