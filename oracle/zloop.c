/* oracle/zloop.c -- TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Plain-C restatement of the reference's four-deep numba loops, same loop order, same
 * float32 arithmetic and the same SEQUENTIAL float32 accumulation:
 *   zl_zigap_z  <- ZIGaP.compute_Z_q_expectations, oriana/models/zigap.py:79-95
 *   zl_gap_z    <- GaP.compute_Z_q_expectations,   oriana/models/gap.py:67-80
 *   zl_sparse_z <- SparseZIGaP.compute_Z_q_expectations, oriana/models/sparse_zigap.py:100-116
 * All arrays are C-contiguous float32 (the numba eager signature, zigap.py:79).
 * `quirk` != 0 reproduces zigap.py:94 (D_hat[i, k]); 0 uses D_hat[i, j] (sparse_zigap.py:115).
 * The third output (DZ_exp_logsum_hat, zigap.py:95; never read by ZIGaP) is filled when Z3 != NULL.
 * Build: make -C oracle   (gcc -O2, no -ffast-math: numba 0.41 compiled without fastmath).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define MAXK 256

void zl_zigap_z(float *Zi, float *Zj, float *Z3, const float *lU, const float *lV, const float *D,
                const float *X, long n, long p, long K, int quirk)
{
    float e[MAXK], ls[MAXK];
    if (Z3) memset(Z3, 0, sizeof(float) * (size_t)(p * K));
    memset(Zi, 0, sizeof(float) * (size_t)(n * K));
    memset(Zj, 0, sizeof(float) * (size_t)(p * K));
    for (long i = 0; i < n; ++i)
        for (long j = 0; j < p; ++j) {
            float den = 0.f;
            for (long k = 0; k < K; ++k) { ls[k] = lU[i * K + k] + lV[j * K + k]; e[k] = expf(ls[k]); den += e[k]; }
            if (!(den > 0.f)) den = 1.f;
            for (long k = 0; k < K; ++k) {
                float t = X[i * p + j] * e[k] / den;
                Zi[i * K + k] += D[i * p + j] * t;
                Zj[j * K + k] += (quirk ? D[i * p + k] : D[i * p + j]) * t;
                if (Z3) Z3[j * K + k] += D[i * p + j] * t * ls[k];
            }
        }
}

void zl_gap_z(float *Zi, float *Zj, const float *lU, const float *lV, const float *X,
              long n, long p, long K)
{
    float e[MAXK];
    memset(Zi, 0, sizeof(float) * (size_t)(n * K));
    memset(Zj, 0, sizeof(float) * (size_t)(p * K));
    for (long i = 0; i < n; ++i)
        for (long j = 0; j < p; ++j) {
            float den = 0.f;
            for (long k = 0; k < K; ++k) { e[k] = expf(lU[i * K + k] + lV[j * K + k]); den += e[k]; }
            if (!(den > 0.f)) den = 1.f;
            for (long k = 0; k < K; ++k) {
                float t = X[i * p + j] * e[k] / den;
                Zj[j * K + k] += t;
                Zi[i * K + k] += t;
            }
        }
}

void zl_sparse_z(float *DSZ, float *DZ, float *DZl, const float *lU, const float *lV, const float *St, const float *Sh,
                 const float *D, const float *X, long n, long p, long K)
{
    float e[MAXK], ls[MAXK];
    memset(DSZ, 0, sizeof(float) * (size_t)(n * K));
    memset(DZ, 0, sizeof(float) * (size_t)(p * K));
    memset(DZl, 0, sizeof(float) * (size_t)(p * K));
    for (long i = 0; i < n; ++i)
        for (long j = 0; j < p; ++j) {
            float den = 0.f;
            for (long k = 0; k < K; ++k) {
                ls[k] = lU[i * K + k] + lV[j * K + k];
                e[k] = expf(ls[k]) * St[j * K + k];
                den += e[k];
            }
            if (!(den > 0.f)) den = 1.f;
            for (long k = 0; k < K; ++k) {
                float t = X[i * p + j] * e[k] / den;
                DSZ[i * K + k] += D[i * p + j] * Sh[j * K + k] * t;
                DZ[j * K + k] += D[i * p + j] * t;
                DZl[j * K + k] += D[i * p + j] * t * ls[k];
            }
        }
}
