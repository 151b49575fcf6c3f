// special.cuh -- device special functions of the CAVI path.
//
// Replaces oriana/utils.py:9-51 of the reference (logit, sigmoid, digamma = scipy.special.digamma,
// digamma_prime = scipy.special.polygamma(1, .), inverse_digamma = Minka start + 5 Newton steps).
// digamma / trigamma: upward recurrence to x >= 10, then the asymptotic (Bernoulli) series;
// absolute error < 1e-13 in double for x in [1e-15, 1e300].
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ori {

constexpr double kDigammaOne = -0.57721566490153286061;  // psi(1), utils.py:48

__host__ __device__ inline double digamma_f64(double x) {
    double r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    const double i = 1.0 / x, i2 = i * i;
    // ln x - 1/2x - sum B_2n / (2n x^2n)
    double s = i2 * (1.0 / 12.0 - i2 * (1.0 / 120.0 - i2 * (1.0 / 252.0 - i2 * (1.0 / 240.0
             - i2 * (1.0 / 132.0 - i2 * (691.0 / 32760.0 - i2 * (1.0 / 12.0)))))));
    return r + log(x) - 0.5 * i - s;
}

__host__ __device__ inline double trigamma_f64(double x) {
    double r = 0.0;
    while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
    const double i = 1.0 / x, i2 = i * i;
    // 1/x + 1/2x^2 + sum B_2n / x^(2n+1)
    double s = i * (1.0 + i * 0.5 + i2 * (1.0 / 6.0 - i2 * (1.0 / 30.0 - i2 * (1.0 / 42.0
             - i2 * (1.0 / 30.0 - i2 * (5.0 / 66.0 - i2 * (691.0 / 2730.0 - i2 * (7.0 / 6.0))))))));
    return r + s;
}

// utils.py:39-51
__host__ __device__ inline double inverse_digamma_f64(double y) {
    double x = (y >= -2.22) ? exp(y) + 0.5 : -1.0 / (y - kDigammaOne);
#pragma unroll 1
    for (int it = 0; it < 5; ++it) x -= (digamma_f64(x) - y) / trigamma_f64(x);
    return x;
}

// utils.py:9-11
__host__ __device__ inline double logit_f64(double x) {
    x = fmin(fmax(x, 1e-15), 1.0 - 1e-15);
    return log(x / (1.0 - x));
}

// utils.py:14-15
__host__ __device__ inline double sigmoid_f64(double x) { return 1.0 / (1.0 + exp(-x)); }

// max(1e-15, nan_to_num(x)) -- zigap.py:117-118 etc.  (NaN -> 0 -> 1e-15, +inf -> DBL_MAX)
__host__ __device__ inline double clamp_param_f64(double x) {
    if (x != x) x = 0.0;
    if (x > 1.7976931348623157e308) x = 1.7976931348623157e308;
    return fmax(1e-15, x);
}
// float32 storage of the same clamp: +inf -> FLT_MAX
__host__ __device__ inline float clamp_param_f32(float x) {
    if (x != x) x = 0.f;
    if (x > 3.402823466e38f) x = 3.402823466e38f;
    return fmaxf(1e-15f, x);
}


// exp(E[log .]) is kept per row (cell or gene) as exp(Elog - max_k Elog + EXP_CENTER): the multinomial step only uses
// ratios exp(lU_ik + lV_jk) / sum_k'(...) (zigap.py:86-92), which do not change when a row of either factor is
// rescaled, and the centred operands stay inside float32 where exp(lU) * exp(lV) would under- or overflow although
// exp(lU + lV) -- what the reference evaluates -- does not (NMF-initialised factors, base.py:38-40, live there).
// Range plan: operands in [e^-64, e^40.2] (components more than e^-104 below their row's largest are flushed to 0: the
// reference's float32 exp(lU + lV) underflows below -103.3, so with log-expectations of order 1 on the other side
// those terms are 0 there too); denominators <= K e^80; accumulators of R * operand <= (sum of X) * e^64.
// The scale is a power of two, 2^58 = e^40.2: the largest component of every row becomes exactly 2^58, which tf32
// represents exactly -- a scale like e^40 would give every row's leading operand the SAME tf32 rounding error, a bias
// the sums over cells / genes do not average out (measured: +50 % error of the tensor path against the fp32 path).
constexpr int EXP_CENTER_LOG2 = 58;
constexpr double EXP_CENTER = 58.0 * 0.69314718055994530942;     // in natural-log units (ELBO bookkeeping)
constexpr float EXP_FLUSH = -104.f;
// rows whose largest log-expectation is below this carry no representable ratio information in float32 (the log itself
// has lost its fractional bits): all their components are treated as 0, as the reference's float32 exp would make them
constexpr float EXP_DEAD = -1e4f;

__host__ __device__ inline float centred_exp_f32(float elog, float row_max) {
    const float rel = elog - row_max;
    return (row_max > EXP_DEAD && rel >= EXP_FLUSH) ? (float)ldexp(exp((double)rel), EXP_CENTER_LOG2) : 0.f;
}

// Float32-underflow emulation (zigap.py:86-90, gap.py:73-76): the reference evaluates exp(lU_ik + lV_jk) in float32,
// which rounds to 0 below 2^-150 (log-sum <= -103.972); an entry whose every term does so keeps den = 0 -> 1 and hands
// its count to no component.  In centred operands term_k = eU_ik * eV_jk = exp(lU + lV - m_i - m_j) 2^116, so
//     term_k underflows in the reference  <=>  eU_ik * eV_jk <= thr_i * thr_j,   thr = 2^-17 exp(-m)
// with m the row's largest log-expectation.  Saturates at 3e38 (rows that far down pair with nothing real) and is 3e38 for
// the all-zero rows of EXP_DEAD.
__host__ __device__ inline float underflow_thr_f32(float row_max) {
    if (!(row_max > EXP_DEAD)) return 3.0e38f;
    const double t = ldexp(exp(-(double)row_max), -17);
    return t > 3.0e38 ? 3.0e38f : (float)t;
}
// an entry can only be touched by the rule when its denominator is within 2^24 of the threshold (terms 2^24 below the
// denominator do not move a float32 sum)
constexpr float UFL_NEAR = 16777216.f;

}  // namespace ori
