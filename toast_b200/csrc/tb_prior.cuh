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

// ------------------------------------------------------------------------------------------------
// Partitioned form of the same solve.  NOT wired into a kernel yet: the cores below are checked on
// the host against scipy (tests/test_offset_prior.py, looped as the six launches would run them);
// k_prior_banded still solves one segment per thread.  A segment is cut
// into chunks of m >= w - 1 rows.  Forward substitution of chunk [s, e) only needs the w - 1
// values before s, so
//     y[s:e] = u + Gf t,   u = the chunk solved with zeros before it (chunks in parallel),
//                          t = y[s-(w-1) : s],  Gf = response of the chunk to unit entries of t
// (Gf is precomputed once per factor); the tails t_p follow from a short recurrence over chunks,
// and the correction u + Gf t is again parallel over rows.  The back substitution is the mirror
// image with the w - 1 values AFTER the chunk.  Same arithmetic as the sequential solve up to
// rounding (the chunk-local triangular systems are the diagonal blocks of L).
// ------------------------------------------------------------------------------------------------

// rows [s, e) of L y = rhs, ignoring everything before row s.  `y` is indexed from row 0.
TBP_HD void fwd_chunk(const double *ab, int64_t w, int64_t n, int64_t s, int64_t e,
                      const double *rhs, double *y) {
    for (int64_t j = s; j < e; ++j) {
        double acc = rhs[j];
        const int64_t kmax = (j - s) < w - 1 ? (j - s) : w - 1;
        for (int64_t k = 1; k <= kmax; ++k) acc -= ab[k * n + (j - k)] * y[j - k];
        y[j] = acc / ab[j];
    }
}

// rows [s, e) of L^T x = rhs, ignoring everything from row e on.  x may alias rhs.
TBP_HD void bwd_chunk(const double *ab, int64_t w, int64_t n, int64_t s, int64_t e,
                      const double *rhs, double *x) {
    for (int64_t j = e - 1; j >= s; --j) {
        double acc = rhs[j];
        const int64_t kmax = (e - 1 - j) < w - 1 ? (e - 1 - j) : w - 1;
        for (int64_t k = 1; k <= kmax; ++k) acc -= ab[k * n + j] * x[j + k];
        x[j] = acc / ab[j];
    }
}

// Gf[(j - s) * (w - 1) + c], j in [s, e): response of the chunk's forward substitution to a unit
// value at row s - (w - 1) + c (c = 0 .. w-2) with a zero right-hand side.  Needs s >= w - 1.
TBP_HD void fwd_response(const double *ab, int64_t w, int64_t n, int64_t s, int64_t e, double *Gf) {
    const int64_t q = w - 1;
    for (int64_t c = 0; c < q; ++c) {
        const int64_t src = s - q + c; // the row carrying the unit value
        for (int64_t j = s; j < e; ++j) {
            double acc = 0.0;
            // coupling to the unit entry: L[j][src] = ab[j - src][src] when 1 <= j - src <= w - 1
            const int64_t d = j - src;
            if (d >= 1 && d <= q) acc -= ab[d * n + src];
            const int64_t kmax = (j - s) < q ? (j - s) : q;
            for (int64_t k = 1; k <= kmax; ++k)
                acc -= ab[k * n + (j - k)] * Gf[(j - k - s) * q + c];
            Gf[(j - s) * q + c] = acc / ab[j];
        }
    }
}

// Gb[(j - s) * (w - 1) + c], j in [s, e): response of the chunk's back substitution to a unit
// value at row e + c (c = 0 .. w-2; rows >= n do not exist and give zero columns).
TBP_HD void bwd_response(const double *ab, int64_t w, int64_t n, int64_t s, int64_t e, double *Gb) {
    const int64_t q = w - 1;
    for (int64_t c = 0; c < q; ++c) {
        const int64_t src = e + c;
        for (int64_t j = e - 1; j >= s; --j) {
            double acc = 0.0;
            // coupling to the unit entry: L[src][j] = ab[src - j][j] when 1 <= src - j <= w - 1
            const int64_t d = src - j;
            if (src < n && d >= 1 && d <= q) acc -= ab[d * n + j];
            const int64_t kmax = (e - 1 - j) < q ? (e - 1 - j) : q;
            for (int64_t k = 1; k <= kmax; ++k)
                acc -= ab[k * n + j] * Gb[(j + k - s) * q + c];
            Gb[(j - s) * q + c] = acc / ab[j];
        }
    }
}

// value of row j (in chunk [s, e)) after adding the response to the boundary vector t[0..w-2]
TBP_HD double chunk_correct(const double *G, int64_t w, int64_t s, int64_t j, double base,
                            const double *t) {
    const int64_t q = w - 1;
    const double *g = G + (j - s) * q;
    double acc = base;
    for (int64_t c = 0; c < q; ++c) acc += g[c] * t[c];
    return acc;
}

} // namespace tbp
