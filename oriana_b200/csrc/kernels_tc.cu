// kernels_tc.cu -- the two X-streaming passes of the CAVI iteration on the sm_100a tensor path:
// TMA-staged tiles, tcgen05.mma kind::tf32 with fp32 accumulators in TMEM, the ratio / dropout posterior
// computed between the two groups of contractions on the TMEM-resident tile (A operand from TMEM).
//
// Per tile of 128 "own" x 64 "sweep" entries (own = cells in the row pass, genes in the gene pass):
//   S:  den = eO . eS^T     (3xTF32: hi.hi + hi.lo + lo.hi)          zigap.py:86-90
//       uv  = Oh . Sh^T     (3xTF32)                                 zigap.py:131 (U_hat V_hat^T)
//   E:  R = X / den ; D = X != 0 ? 1 : max(sigmoid(lp - uv), floor)  zigap.py:91-92, :131-136
//       (written back over den / uv in TMEM, rounded to tf32 to nearest)
//   P:  acc1 += R . S1 ; acc2 += D . S2                              zigap.py:93-94 / :116, :124
// where S1, S2 are the transposed (K-major) copies of exp(E log .) and of the matching U_hat / V_hat.
// MN-major tf32 operands would need the SWIZZLE_128B_ATOM_32B smem layout, which no K-major operand accepts,
// hence the transposed copies (prepared by k_tc_prep_*) and K-major descriptors everywhere
// (scripts/tc_probe.cu checks every descriptor form used here against the CPU).
//
// Warp roles (320 threads, 1 CTA per SM, persistent over work items):
//   warp 0    TMA producer            (one lane)
//   warp 1    MMA issuer + TMEM owner (one lane issues)
//   warps 2-9 element-wise stage + epilogue; warp w owns TMEM lanes 32*(w%4).. and half of the 64 columns
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ori {
using namespace tc;

constexpr int TC_OWN = 128;
constexpr int TC_SW = 64;
constexpr int TC_KP_CONST = 32;     // latent dimension of the tensor path (K <= 32, zero padded)
constexpr int TC_EW_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EW_WARPS;

constexpr uint32_t RES_BYTES = 4 * 16384;          // own-side K-major operands: eO_hi, eO_lo, Oh_hi, Oh_lo [128 x 32]
constexpr uint32_t ST_K = 0;                       // sweep-side K-major operands: 4 x [64 x 32]
constexpr uint32_t ST_T = 32768;                   // sweep-side transposed operands: 2 arrays x 2 chunks [32 x 32]
constexpr uint32_t ST_X = 49152;                   // X tile, 32 KB
constexpr uint32_t STAGE_BYTES = 81920;
constexpr uint32_t LP_OFF = RES_BYTES + 2 * STAGE_BYTES;   // per stage: lp2[64] | floor[64]
constexpr uint32_t BAR_OFF = LP_OFF + 1024;
constexpr uint32_t TC_SMEM_BYTES = BAR_OFF + 256 + 1024;   // + barriers + alignment slack

enum { B_RES_FULL = 0, B_RES_EMPTY = 1, B_FULL = 2, B_EMPTY = 4, B_SREADY = 6, B_PREADY = 8, B_ACC_READY = 10,
       B_ACC_FREE = 11, NBARS = 12 };

constexpr uint32_t TM_STAGE = 128;   // TMEM columns per stage: den/R [0,64) | uv/D [64,128)
constexpr uint32_t TM_ACC1 = 256;    // 32 columns
constexpr uint32_t TM_ACC2 = 288;    // 32 columns
constexpr uint32_t TM_COLS = 512;

struct TcMaps { CUtensorMap ownK, swK, swT, X; };

struct TcArgs {
    long long own_total, sw_total;   // valid extents (cells / genes)
    long long own_pad, sw_pad;       // padded extents (multiples of 128): q-th operand array starts at row q*pad
    int n_own_tiles, n_chunks, tiles_per_chunk, n_sw_tiles;
    const float* lp2w;               // [genes_pad] logit(pi) * log2(e); -inf: D_hat = (X>0)
    const float* flw;                // [genes_pad] floor (1e-10 where pi <= 0)
    float* acc1;                     // [own_total x 32]  sum_sweep R  * S1
    float* acc2;                     // [own_total x 32]  sum_sweep D  * S2
    double* colsum;                  // gene pass: [genes] += sum_i D_hat
    double* part64;                  // gene pass: ELBO partial sums
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

template <bool GENES, bool DROPOUT, bool ELBO>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_pass(const __grid_constant__ TcMaps maps, const TcArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + BAR_OFF);
    uint32_t* tmem_slot = (uint32_t*)(bars + NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[B_RES_FULL], 1);
        mbar_init(&bars[B_RES_EMPTY], 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[B_FULL + s], 1);
            mbar_init(&bars[B_EMPTY + s], 1 + TC_EW_WARPS);
            mbar_init(&bars[B_SREADY + s], 1);
            mbar_init(&bars[B_PREADY + s], TC_EW_WARPS);
        }
        mbar_init(&bars[B_ACC_READY], 1);
        mbar_init(&bars[B_ACC_FREE], TC_EW_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_items = a.n_own_tiles * a.n_chunks;
    constexpr int NQ = DROPOUT ? 4 : 2;       // K-major operand arrays in use
    constexpr int NT = DROPOUT ? 2 : 1;       // transposed operand arrays in use

    if (warp == 0) {
        // ============================================ TMA producer ============================================
        if (lane == 0) {
            tma_prefetch_desc(&maps.ownK); tma_prefetch_desc(&maps.swK); tma_prefetch_desc(&maps.swT); tma_prefetch_desc(&maps.X);
            uint32_t it = 0;
            int li = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
                const int own_tile = item / a.n_chunks, chunk = item % a.n_chunks;
                const int own0 = own_tile * TC_OWN;
                const int t_begin = chunk * a.tiles_per_chunk;
                const int t_end = min(t_begin + a.tiles_per_chunk, a.n_sw_tiles);
                mbar_wait(&bars[B_RES_EMPTY], (li & 1) ^ 1, 10);
                mbar_expect_tx(&bars[B_RES_FULL], NQ * 16384);
                for (int q = 0; q < NQ; ++q)
                    tma_load_2d(smem + q * 16384, &maps.ownK, &bars[B_RES_FULL], 0, (int)(q * a.own_pad + own0));
                for (int t = t_begin; t < t_end; ++t, ++it) {
                    const int s = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    mbar_wait(&bars[B_EMPTY + s], ph ^ 1, 11);
                    uint8_t* st = smem + RES_BYTES + s * STAGE_BYTES;
                    uint64_t* bar = &bars[B_FULL + s];
                    const int sw0 = t * TC_SW;
                    uint32_t bytes = NQ * 8192 + NT * 8192 + 32768;
                    if (!GENES && DROPOUT) bytes += 512;
                    mbar_expect_tx(bar, bytes);
                    for (int q = 0; q < NQ; ++q)
                        tma_load_2d(st + ST_K + q * 8192, &maps.swK, bar, 0, (int)(q * a.sw_pad + sw0));
                    for (int q = 0; q < NT; ++q)
                        for (int c = 0; c < 2; ++c)
                            tma_load_2d(st + ST_T + q * 8192 + c * 4096, &maps.swT, bar, sw0 + 32 * c, q * 32);
                    if (!GENES) {
                        for (int c = 0; c < 2; ++c) tma_load_2d(st + ST_X + c * 16384, &maps.X, bar, sw0 + 32 * c, own0);
                        if (DROPOUT) {
                            bulk_load(smem + LP_OFF + s * 512, a.lp2w + sw0, 256, bar);
                            bulk_load(smem + LP_OFF + s * 512 + 256, a.flw + sw0, 256, bar);
                        }
                    } else {
                        for (int c = 0; c < 4; ++c) tma_load_2d(st + ST_X + c * 8192, &maps.X, bar, own0 + 32 * c, sw0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================================ MMA issuer ===============================================
        if (lane == 0) {
            constexpr uint32_t idescS = make_idesc_tf32(TC_OWN, TC_SW, false, false);
            constexpr uint32_t idescP = make_idesc_tf32(TC_OWN, TC_KP_CONST, false, false);
            const uint32_t res = smem_u32(smem);
            auto issue_P = [&](int s, uint32_t ph, bool first, bool last, int li) {
                mbar_wait(&bars[B_PREADY + s], ph, 20);
                if (first) mbar_wait(&bars[B_ACC_FREE], (li & 1) ^ 1, 21);
                tc_fence_after();
                const uint32_t st = res + RES_BYTES + s * STAGE_BYTES + ST_T;
                for (int ks = 0; ks < 8; ++ks)
                    mma_tf32_ts(tmem + TM_ACC1, tmem + s * TM_STAGE + ks * 8,
                                make_smem_desc(st + (ks >> 2) * 4096 + (ks & 3) * 32, 16, 1024), idescP, !(first && ks == 0));
                if (DROPOUT)
                    for (int ks = 0; ks < 8; ++ks)
                        mma_tf32_ts(tmem + TM_ACC2, tmem + s * TM_STAGE + 64 + ks * 8,
                                    make_smem_desc(st + 8192 + (ks >> 2) * 4096 + (ks & 3) * 32, 16, 1024), idescP,
                                    !(first && ks == 0));
                tc_commit(&bars[B_EMPTY + s]);
                if (last) tc_commit(&bars[B_ACC_READY]);
            };
            uint32_t it = 0;
            int li = 0;
            bool have_prev = false, prev_first = false, prev_last = false;
            int prev_s = 0, prev_li = 0;
            uint32_t prev_ph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
                const int chunk = item % a.n_chunks;
                const int t_begin = chunk * a.tiles_per_chunk;
                const int t_end = min(t_begin + a.tiles_per_chunk, a.n_sw_tiles);
                for (int t = t_begin; t < t_end; ++t, ++it) {
                    const int s = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    if (t == t_begin) mbar_wait(&bars[B_RES_FULL], li & 1, 22);
                    mbar_wait(&bars[B_FULL + s], ph, 23);
                    tc_fence_after();
                    const uint32_t st = res + RES_BYTES + s * STAGE_BYTES + ST_K;
                    // den (and uv) of this tile: three tf32 products per contraction
                    auto chain = [&](uint32_t d, int qa, int qb, bool first) {
                        for (int kk = 0; kk < 4; ++kk)
                            mma_tf32_ss(d, make_smem_desc(res + qa * 16384 + kk * 32, 16, 1024),
                                        make_smem_desc(st + qb * 8192 + kk * 32, 16, 1024), idescS, !(first && kk == 0));
                    };
                    chain(tmem + s * TM_STAGE, 0, 0, true);
                    chain(tmem + s * TM_STAGE, 0, 1, false);
                    chain(tmem + s * TM_STAGE, 1, 0, false);
                    if (DROPOUT) {
                        chain(tmem + s * TM_STAGE + 64, 2, 2, true);
                        chain(tmem + s * TM_STAGE + 64, 2, 3, false);
                        chain(tmem + s * TM_STAGE + 64, 3, 2, false);
                    }
                    tc_commit(&bars[B_SREADY + s]);
                    if (t == t_end - 1) tc_commit(&bars[B_RES_EMPTY]);
                    if (have_prev) issue_P(prev_s, prev_ph, prev_first, prev_last, prev_li);
                    have_prev = true; prev_s = s; prev_ph = ph; prev_first = (t == t_begin); prev_last = (t == t_end - 1);
                    prev_li = li;
                }
            }
            if (have_prev) issue_P(prev_s, prev_ph, prev_first, prev_last, prev_li);
        }
    } else {
        // ======================================= element-wise stage + epilogue =================================
        const int ew = warp - 2;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
        const int half = ew >> 2;                     // which 32 of the 64 tile columns
        const int lrow = quarter * 32 + lane;         // own index inside the tile = TMEM lane
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        uint32_t it = 0;
        int li = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
            const int own_tile = item / a.n_chunks, chunk = item % a.n_chunks;
            const long long own_idx = (long long)own_tile * TC_OWN + lrow;
            const bool own_ok = own_idx < a.own_total;
            const int t_begin = chunk * a.tiles_per_chunk;
            const int t_end = min(t_begin + a.tiles_per_chunk, a.n_sw_tiles);
            float lp2j = 0.f, flj = 0.f;
            if (GENES && DROPOUT) { lp2j = a.lp2w[own_idx]; flj = a.flw[own_idx]; }   // padded arrays: always in range
            float cs = 0.f;
            double acc_xl = 0.0, acc_ent = 0.0;
            for (int t = t_begin; t < t_end; ++t, ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&bars[B_FULL + s], ph, 30);      // X tile (and lp) visible to this thread
                mbar_wait(&bars[B_SREADY + s], ph, 31);    // den / uv complete in TMEM
                tc_fence_after();
                const uint8_t* Xs = smem + RES_BYTES + s * STAGE_BYTES + ST_X;
                const float* lps = (const float*)(smem + LP_OFF + s * 512);
                const int valid = (int)min((long long)TC_SW, a.sw_total - (long long)t * TC_SW);   // gene pass: real cells
                float t_xl = 0.f, t_ent = 0.f;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int c0 = half * 32 + g * 16;
                    uint32_t den_r[16], uv_r[16];
                    tmem_ld16(tlane + s * TM_STAGE + c0, den_r);
                    if (DROPOUT) tmem_ld16(tlane + s * TM_STAGE + 64 + c0, uv_r);
                    float x[16], lp2[16], fl[16];
                    if (!GENES) {
                        const uint8_t* base = Xs + (c0 >> 5) * 16384 + lrow * 128;
                        const int cb = (c0 & 31) >> 2;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 v = *reinterpret_cast<const float4*>(base + (((cb + q) ^ (lrow & 7)) << 4));
                            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                        }
                        if (DROPOUT) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 l = *reinterpret_cast<const float4*>(lps + c0 + 4 * q);
                                const float4 f = *reinterpret_cast<const float4*>(lps + 64 + c0 + 4 * q);
                                lp2[4 * q] = l.x; lp2[4 * q + 1] = l.y; lp2[4 * q + 2] = l.z; lp2[4 * q + 3] = l.w;
                                fl[4 * q] = f.x; fl[4 * q + 1] = f.y; fl[4 * q + 2] = f.z; fl[4 * q + 3] = f.w;
                            }
                        }
                    } else {
                        const uint8_t* base = Xs + quarter * 8192 + (lane & 3) * 4;
                        const int jc = lane >> 2;
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int i = c0 + e;
                            x[e] = *reinterpret_cast<const float*>(base + i * 128 + ((jc ^ (i & 7)) << 4));
                        }
                    }
                    tmem_wait_ld();
                    uint32_t R_r[16], D_r[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float den = __uint_as_float(den_r[e]);
                        const bool nz = x[e] != 0.f;
                        const float dg = den > 0.f ? den : 1.f;                       // zigap.py:90
                        float tt = dg, e2 = 0.f;
                        if (DROPOUT) {
                            const float uv = __uint_as_float(uv_r[e]);
                            e2 = fminf(fmaxf(fmaf(uv, LOG2E, -(GENES ? lp2j : lp2[e])), -126.f), 127.f);
                            tt = nz ? dg : 1.f + ex2_approx(e2);
                        }
                        const float r = rcp_approx(tt);
                        const float R = x[e] * r;                                     // 0 where X == 0
                        R_r[e] = __float_as_uint(to_tf32_rna(R));
                        if (DROPOUT) {
                            const float D = fmaxf(nz ? 1.f : r, GENES ? flj : fl[e]);  // zigap.py:133-136
                            D_r[e] = __float_as_uint(to_tf32_rna(D));
                            if (GENES) {
                                const bool live = (c0 + e) < valid;
                                cs += live ? D : 0.f;
                                if (ELBO && live) {
                                    const float l2 = lg2_approx(tt);
                                    t_xl = fmaf(x[e], l2, t_xl);
                                    t_ent += (nz ? 0.f : l2) - (1.f - D) * e2;
                                }
                            }
                        } else if (GENES && ELBO) {
                            if ((c0 + e) < valid) t_xl = fmaf(x[e], lg2_approx(tt), t_xl);
                        }
                    }
                    tmem_st16(tlane + s * TM_STAGE + c0, R_r);
                    if (DROPOUT) tmem_st16(tlane + s * TM_STAGE + 64 + c0, D_r);
                }
                if (GENES && ELBO) { acc_xl += (double)t_xl; acc_ent += (double)t_ent; }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&bars[B_PREADY + s]); mbar_arrive(&bars[B_EMPTY + s]); }
            }
            // ---- epilogue of the work item: accumulators -> global (atomics: items of one own-tile may be split)
            mbar_wait(&bars[B_ACC_READY], li & 1, 32);
            tc_fence_after();
            if (half == 0 || DROPOUT) {
                float* out = (half == 0 ? a.acc1 : a.acc2) + own_idx * 32;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t v[16];
                    tmem_ld16(tlane + (half == 0 ? TM_ACC1 : TM_ACC2) + g * 16, v);
                    tmem_wait_ld();
                    if (own_ok) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) atomicAdd(out + g * 16 + e, __uint_as_float(v[e]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_ACC_FREE]);
            if (GENES) {
                if (DROPOUT && own_ok) atomicAdd(a.colsum + own_idx, (double)cs);
                if (ELBO) {
                    if (!own_ok) { acc_xl = 0.0; acc_ent = 0.0; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        acc_xl += __shfl_xor_sync(0xffffffffu, acc_xl, o);
                        acc_ent += __shfl_xor_sync(0xffffffffu, acc_ent, o);
                    }
                    if (lane == 0) {
                        atomicAdd(a.part64 + R64_XLOGDEN, acc_xl * (double)LN2);
                        if (DROPOUT) atomicAdd(a.part64 + R64_ENT, acc_ent * (double)LN2);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TM_COLS);
}

// ---- operand preparation -------------------------------------------------------------------------------------
// K-major operand arrays: out[q][pad][32], q = 0: hi(e) 1: lo(e) 2: hi(E) 3: lo(E); hi = tf32 round-to-nearest
// (the tensor core truncates fp32 operands to tf32: scripts/tc_probe.cu T4), lo = x - hi.  Pad rows are zero.
__global__ void __launch_bounds__(256)
k_tc_prep_K(const float* __restrict__ e, const float* __restrict__ E, float* __restrict__ out, long long n, long long pad)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pad * 32) return;
    const bool ok = idx < n * 32;
    const float v = ok ? e[idx] : 0.f;
    const float hv = to_tf32_rna(v);
    out[idx] = hv;
    out[pad * 32 + idx] = v - hv;
    if (E) {
        const float w = ok ? E[idx] : 0.f;
        const float hw = to_tf32_rna(w);
        out[2 * pad * 32 + idx] = hw;
        out[3 * pad * 32 + idx] = w - hw;
    }
}
// transposed operand: out[k][i] = tf32(src[i][k]),  src [n x 32], out [32 x pad]
__global__ void __launch_bounds__(256)
k_tc_prep_T(const float* __restrict__ src, float* __restrict__ out, long long n, long long pad)
{
    __shared__ float tile[32][33];
    const long long i0 = (long long)blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const long long i = i0 + r;
        tile[r][threadIdx.x] = i < n ? src[i * 32 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += 8) {
        const long long i = i0 + threadIdx.x;
        if (i < pad) out[(long long)k * pad + i] = to_tf32_rna(tile[threadIdx.x][k]);
    }
}
__global__ void k_tc_prep_lp(const float* __restrict__ lp, const float* __restrict__ fl, float* __restrict__ lp2w,
                             float* __restrict__ flw, int p, int pad)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= pad) return;
    lp2w[j] = j < p ? lp[j] * LOG2E : -INFINITY;
    flw[j] = j < p ? fl[j] : 0.f;
}

// ---- host side ---------------------------------------------------------------------------------------------------
static long long pad128(long long v) { return (v + 127) / 128 * 128; }

long long tc_workspace_floats(long long n_rows, int p) {
    const long long np = pad128(n_rows), pp = pad128(p);
    return 4 * np * 32 + 2 * 32 * np + 4 * pp * 32 + 2 * 32 * pp + 2 * pp;
}

struct TcWs { float *rowK, *rowT, *geneK, *geneT, *lp2w, *flw; long long np, pp; };
static TcWs tc_carve(const ori_problem_t* P) {
    TcWs w;
    w.np = pad128(P->n_rows); w.pp = pad128(P->p);
    float* f = P->tc_ws;
    w.rowK = f; f += 4 * w.np * 32;
    w.rowT = f; f += 2 * 32 * w.np;
    w.geneK = f; f += 4 * w.pp * 32;
    w.geneT = f; f += 2 * 32 * w.pp;
    w.lp2w = f; f += w.pp;
    w.flw = f;
    return w;
}

bool tc_eligible(const ori_problem_t* P) {
    return P->tc_ws != nullptr && P->KP == 32 && !(P->flags & ORI_F_NO_TENSOR) && P->n_rows > 0 &&
           P->tc_ws_floats >= tc_workspace_floats(P->n_rows, P->p) && get_encode_fn() != nullptr;
}

static int num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

// gene-side operands (+ padded logit(pi)); run once per iteration before the row pass
int launch_tc_prep_genes(const ori_problem_t* P, cudaStream_t st) {
    const TcWs w = tc_carve(P);
    const bool drop = P->flags & ORI_F_DROPOUT;
    k_tc_prep_K<<<cdiv(w.pp * 32, 256), 256, 0, st>>>(P->eV, drop ? P->V_hat : nullptr, w.geneK, P->p, w.pp);
    k_tc_prep_T<<<cdiv(w.pp, 32), dim3(32, 8), 0, st>>>(P->eV, w.geneT, P->p, w.pp);
    if (drop) {
        k_tc_prep_T<<<cdiv(w.pp, 32), dim3(32, 8), 0, st>>>(P->V_hat, w.geneT + 32 * w.pp, P->p, w.pp);
        k_tc_prep_lp<<<cdiv(w.pp, 256), 256, 0, st>>>(P->lp, P->pfloor, w.lp2w, w.flw, P->p, (int)w.pp);
    }
    return check_launch("k_tc_prep(genes)");
}
// row-side operands of generation g (before the row pass)
int launch_tc_prep_rows(const ori_problem_t* P, int g, cudaStream_t st) {
    const TcWs w = tc_carve(P);
    const bool drop = P->flags & ORI_F_DROPOUT;
    k_tc_prep_K<<<cdiv(w.np * 32, 256), 256, 0, st>>>(P->eU[g], drop ? P->U_hat[g] : nullptr, w.rowK, P->n_rows, w.np);
    return check_launch("k_tc_prep(rows)");
}
// transposed row operands for the gene pass: the Zj weight (eU, or eU * D_hat[:, :K] under the quirk) and the
// NEW U_hat (zigap.py:124); run after the U update
int launch_tc_prep_rows_T(const ori_problem_t* P, int g, cudaStream_t st) {
    const TcWs w = tc_carve(P);
    const bool drop = P->flags & ORI_F_DROPOUT;
    const float* wsrc = (P->flags & ORI_F_QUIRK) ? P->eUw : P->eU[g];
    k_tc_prep_T<<<cdiv(w.np, 32), dim3(32, 8), 0, st>>>(wsrc, w.rowT, P->n_rows, w.np);
    if (drop) k_tc_prep_T<<<cdiv(w.np, 32), dim3(32, 8), 0, st>>>(P->U_hat[1 - g], w.rowT + 32 * w.np, P->n_rows, w.np);
    return check_launch("k_tc_prep(rows, transposed)");
}

template <bool GENES>
static int launch_tc_pass(const ori_problem_t* P, cudaStream_t st) {
    const TcWs w = tc_carve(P);
    const bool drop = P->flags & ORI_F_DROPOUT, elbo = P->flags & ORI_F_ELBO;
    TcMaps maps;
    TcArgs a;
    bool ok;
    if (!GENES) {
        ok = make_tmap_f32(&maps.ownK, w.rowK, 4 * w.np, 32, 32, 32, 128) &&
             make_tmap_f32(&maps.swK, w.geneK, 4 * w.pp, 32, 32, 32, 64) &&
             make_tmap_f32(&maps.swT, w.geneT, 64, w.pp, w.pp, 32, 32) &&
             make_tmap_f32(&maps.X, P->X, P->n_rows, P->p, P->ldx, 32, 128);
        a.own_total = P->n_rows; a.sw_total = P->p; a.own_pad = w.np; a.sw_pad = w.pp;
        a.acc1 = P->Zi; a.acc2 = P->a2s;
    } else {
        ok = make_tmap_f32(&maps.ownK, w.geneK, 4 * w.pp, 32, 32, 32, 128) &&
             make_tmap_f32(&maps.swK, w.rowK, 4 * w.np, 32, 32, 32, 64) &&
             make_tmap_f32(&maps.swT, w.rowT, 64, w.np, w.np, 32, 32) &&
             make_tmap_f32(&maps.X, P->X, P->n_rows, P->p, P->ldx, 32, 64);
        a.own_total = P->p; a.sw_total = P->n_rows; a.own_pad = w.pp; a.sw_pad = w.np;
        a.acc1 = P->red32; a.acc2 = P->red32 + (long long)P->p * 32;
    }
    if (!ok) return set_error(ORI_ECUDA, "cuTensorMapEncodeTiled failed");
    a.lp2w = w.lp2w; a.flw = w.flw;
    a.colsum = P->red64; a.part64 = P->red64 + P->p + 2 * P->KP;
    a.n_own_tiles = cdiv(a.own_total, TC_OWN);
    a.n_sw_tiles = cdiv(a.sw_total, TC_SW);
    // split the sweep so that there are enough work items for every SM (and bounded fp32 running sums)
    const int sms = num_sms();
    int chunks = 1;
    while ((long long)a.n_own_tiles * chunks < 4LL * sms && a.n_sw_tiles / (chunks * 2) >= 8) chunks *= 2;
    int tpc = cdiv(a.n_sw_tiles, chunks);
    if (GENES && tpc > 128) tpc = 128;          // gene pass: <= 8192 cells per item (fp32 running sums of the stats)
    a.tiles_per_chunk = tpc;
    a.n_chunks = cdiv(a.n_sw_tiles, tpc);
    const int n_items = a.n_own_tiles * a.n_chunks;
    const int grid = n_items < sms ? n_items : sms;

#define ORI_TC_LAUNCH(D, E)                                                                                     \
    do {                                                                                                        \
        auto kern = k_tc_pass<GENES, D, E>;                                                                     \
        static bool attr_done = false;                                                                          \
        if (!attr_done) {                                                                                       \
            cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES); \
            if (e_ != cudaSuccess) return set_error(ORI_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e_)); \
            attr_done = true;                                                                                   \
        }                                                                                                       \
        kern<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(maps, a);                                                 \
    } while (0)
    if (drop && elbo) ORI_TC_LAUNCH(true, true);
    else if (drop) ORI_TC_LAUNCH(true, false);
    else if (elbo) ORI_TC_LAUNCH(false, true);
    else ORI_TC_LAUNCH(false, false);
#undef ORI_TC_LAUNCH
    return check_launch(GENES ? "k_tc_pass(genes)" : "k_tc_pass(rows)");
}

int launch_pass_rows_tc(const ori_problem_t* P, cudaStream_t st) { return launch_tc_pass<false>(P, st); }
int launch_pass_genes_tc(const ori_problem_t* P, cudaStream_t st) { return launch_tc_pass<true>(P, st); }

}  // namespace ori
