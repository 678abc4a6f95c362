// tb_prior.cuh -- per-element cores of the Offset noise prior / preconditioner kernels, written so
// that the SAME source compiles for the device (tb_prior.cu) and for the host
// (tests/csrc/host_math.cpp checks them against scipy without a GPU).
//
// Reference: templates/offset/offset.py:884-960 (_add_prior: scipy.signal.convolve mode "same")
// and :962-1010 (_apply_precond: scipy.linalg.cho_solve_banded on the lower banded Cholesky
// factor, or a second "same" convolution for the Toeplitz form).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define TBP_HD __host__ __device__ __forceinline__
#else
#define TBP_HD inline
#endif

namespace tbp {

// Element i of convolve(a[0..n), f[0..nf), mode="same"): the full convolution
// c[m] = sum_j f[j] a[m - j] sampled at m = i + (nf - 1) / 2  (scipy _centered on in1's shape).
TBP_HD double conv_same_at(const double *a, int64_t n, const double *f, int64_t nf, int64_t i) {
    const int64_t m = i + (nf - 1) / 2;
    int64_t j0 = m - (n - 1);
    if (j0 < 0) j0 = 0;
    int64_t j1 = m < nf - 1 ? m : nf - 1;
    double acc = 0.0;
    for (int64_t j = j0; j <= j1; ++j) acc += f[j] * a[m - j];
    return acc;
}

// Solve (L L^T) x = b for one segment.  `ab` is the LOWER banded Cholesky factor as
// scipy.linalg.cholesky_banded(lower=True) returns it: ab[k * n + j] = L[j + k][j], k < w.
// x may alias b.  Forward substitution by rows, back substitution by rows of L^T.
TBP_HD void banded_cho_solve(const double *ab, int64_t w, int64_t n, const double *b, double *x) {
    for (int64_t j = 0; j < n; ++j) {
        double s = b[j];
        const int64_t kmax = j < w - 1 ? j : w - 1;
        for (int64_t k = 1; k <= kmax; ++k) s -= ab[k * n + (j - k)] * x[j - k];
        x[j] = s / ab[j];
    }
    for (int64_t j = n - 1; j >= 0; --j) {
        double s = x[j];
        const int64_t kmax = (n - 1 - j) < w - 1 ? (n - 1 - j) : w - 1;
        for (int64_t k = 1; k <= kmax; ++k) s -= ab[k * n + j] * x[j + k];
        x[j] = s / ab[j];
    }
}

} // namespace tbp
