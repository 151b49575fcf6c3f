// common.cuh -- shared declarations of the oriana_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/oriana_b200.h"

namespace ori {

// error plumbing (api.cu)
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what, int n_kernels = 1);   // also counts the kernels launched (ori_kernel_launches)

// slots of ori_problem_t::red64 after the p column sums and the 2K row-side sums
enum Red64Slot {
    R64_XLOGDEN = 0,  // sum_{X>0} X log den
    R64_ENT = 1,      // sum_{X==0} entropy of q(D_ij)
    R64_PUV = 2,      // sum_ij D_hat_ij (U_hat V_hat^T)_ij
    R64_HROW = 3,     // sum_ik entropy of q(U_ik)
    R64_NSLOTS = 8
};
// slots of ori_problem_t::scal (per-model persistent scalars, float64)
enum ScalSlot {
    SC_LGAMX = 0,     // sum_{X>0} lgamma(X+1)      (constant, all ranks)
    SC_NNZ = 1,       // number of non-zero entries (constant, all ranks)
    SC_PENDING = 2,   // K-/p-/nK-sized ELBO terms of the current state
    SC_HGENE = 3,     // sum_jk entropy of q(V_jk) of the current state
    SC_ELBO_LAST = 4, // last finalised ELBO
    SC_ITER = 5,      // number of completed iterations
    SC_NSLOTS = 16
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// kernels_simt.cu
int launch_pass_rows_simt(const ori_problem_t* P, int gen_old, cudaStream_t st);
int launch_pass_genes_simt(const ori_problem_t* P, int gen_old, cudaStream_t st);
int launch_row_update(const ori_problem_t* P, int gen_old, int write_state, cudaStream_t st);
int launch_gene_update(const ori_problem_t* P, int write_state, cudaStream_t st);
int launch_mstep(const ori_problem_t* P, int mode, cudaStream_t st);
int launch_count_stats(const ori_problem_t* P, cudaStream_t st);
int launch_quirk_weights(const ori_problem_t* P, int gen_old, cudaStream_t st);
int launch_dropout_posterior(const ori_problem_t* P, int gen, float* out, long long ldo,
                             long long row0, long long nrows, cudaStream_t st);
int launch_row_sums(const float* X, long long ldx, long long n_rows, int p, float* out, cudaStream_t st);
int launch_col_sums(const float* X, long long ldx, long long n_rows, int p, double* out, cudaStream_t st);
int launch_deviance(const ori_problem_t* P, int gen, const double* pi, const double* col_mean, long long* out_int,
                    double* out_f64, cudaStream_t st);

// kernels_tc.cu
long long tc_workspace_floats(long long n_rows, int p, int KP);

// ORI_F_DETERMINISTIC scratch (ori_problem_t::det_ws), in doubles:
//   [0, DET_FU_BLOCKS * DET_FU_SLOTS)   per-block partial sums of k_factor_update (sum log E | sum E | entropy | D.uv)
//   [.., + pad128(p))                    column sums of D_hat of the second column slice of the tensor gene pass
//   [.., + det_max_items * DET_ITEM_SLOTS)  per-item, per-CTA, per-warp ELBO terms of the tensor gene pass
//   [.., + tickets)                      one int per own tile of the running tensor pass (chunk order of the accumulator adds)
constexpr int DET_FU_BLOCKS = 148 * 8;
constexpr int DET_FU_SLOTS = 2 * 64 + 2;
constexpr int DET_ITEM_SLOTS = 2 * 8 * 2;          // CTAs of a pair x element-wise warps x (x log den, entropy)
//   CUDA-core kernels (problems up to DET_SIMT_MAX entries), from det_simt_offset():
//   [nb x p] column sums per row block | [nb x 2] ELBO terms per row block | [chunks x 3 x p x KP] floats: gene sums per row chunk
constexpr long long DET_SIMT_MAX = 1ll << 26;
long long det_max_items(long long n_rows, int p);
long long det_workspace_doubles(long long n_rows, int p, int KP);
long long det_simt_offset(long long n_rows, int p);
inline bool det_simt_ok(long long n_rows, int p) { return n_rows * (long long)p <= DET_SIMT_MAX; }
// src[n][2] added in index order by one block: *dst0 += sum src[.][0], *dst1 += sum src[.][1]
int launch_det_sum_pairs(const double* src, long long n, double* dst0, double* dst1, cudaStream_t st);
bool tc_eligible(const ori_problem_t* P);
int launch_tc_prep_genes(const ori_problem_t* P, cudaStream_t st);
int launch_tc_prep_rows(const ori_problem_t* P, int gen_old, cudaStream_t st);
int launch_pass_rows_tc(const ori_problem_t* P, int gen_old, cudaStream_t st);
int launch_pass_genes_tc(const ori_problem_t* P, int gen_old, cudaStream_t st);
int launch_tc_prep_rows_logsum(const ori_problem_t* P, int gen_old, cudaStream_t st);
int launch_pass_genes_logsum_tc(const ori_problem_t* P, int gen_old, cudaStream_t st);
int launch_deviance_tc(const ori_problem_t* P, int gen, const double* pi, const double* col_mean, long long* out_int, cudaStream_t st);

}  // namespace ori
