// tb_prior.cuh -- per-element cores of the Offset noise prior / preconditioner kernels, written so
// that the SAME source compiles for the device (tb_prior.cu) and for the host
// (tests/csrc/host_math.cpp checks them against scipy without a GPU).
//
// Reference: templates/offset/offset.py:884-960 (_add_prior: scipy.signal.convolve mode "same")
// and :962-1010 (_apply_precond: scipy.linalg.cho_solve_banded on the lower banded Cholesky
// factor, or a second "same" convolution for the Toeplitz form).
#pragma once

#include <stdint.h>

#include <vector>

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
// Partitioned form of the same solve (k_pb_* in tb_prior.cu; tb_set_option("prior_chunk", m),
// default m = 1024, 0 = one thread per segment).  The code below is checked on the host against
// scipy and the reference's preconditioner (tests/test_offset_prior.py, looped in the order of
// the six launches) and on the device (tests/test_gpu_prior.py).  A segment is cut
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

// ---- the six launches of the partitioned solve, one function per THREAD --------------------------
// (k_pb_* in tb_prior.cu call these with their global thread index; tests/csrc/host_math.cpp
// calls them in loops, so the host check runs the device code path line by line.)
struct PartView {
    int64_t n_seg, n_chunk, qmax;
    const int64_t *seg_start, *seg_len, *p_start, *p_width; // as tb_offset_prior
    const double *factors;
    const int64_t *seg_chunk0; // [n_seg + 1] first global chunk of a segment
    const int64_t *seg_m;      // [n_seg] chunk length (>= width - 1)
    const int64_t *chunk_seg;  // [n_chunk] segment of a chunk
    const int64_t *g_off;      // [n_seg] offset of the segment's response arrays (len x q doubles)
    const double *Gf, *Gb;
    double *tails, *heads;     // [n_chunk * qmax] scratch
};

struct ChunkRef {
    int64_t seg, p, P, s, e, n, w, q, s0;
    const double *ab;
    bool cut;
};

TBP_HD ChunkRef chunk_ref(const PartView &v, int64_t g) {
    ChunkRef c;
    c.seg = v.chunk_seg[g];
    c.p = g - v.seg_chunk0[c.seg];
    c.P = v.seg_chunk0[c.seg + 1] - v.seg_chunk0[c.seg];
    c.n = v.seg_len[c.seg];
    c.w = v.p_width[c.seg];
    c.q = c.w - 1;
    c.s0 = v.seg_start[c.seg];
    const int64_t m = v.seg_m[c.seg];
    c.s = c.p * m;
    c.e = (c.s + m < c.n) ? c.s + m : c.n;
    c.cut = v.p_start[c.seg] < 0;
    c.ab = c.cut ? nullptr : v.factors + v.p_start[c.seg];
    return c;
}

// launch 1 (thread = chunk): chunk-local forward solve, in -> out; cut segments give zeros
TBP_HD void pb_fwd_local(const PartView &v, int64_t g, const double *in, double *out) {
    const ChunkRef c = chunk_ref(v, g);
    if (c.cut) {
        for (int64_t j = c.s; j < c.e; ++j) out[c.s0 + j] = 0.0;
        return;
    }
    fwd_chunk(c.ab, c.w, c.n, c.s, c.e, in + c.s0, out + c.s0);
}

// launch 2 (thread = segment): tails t[p] = y[e_p - q : e_p], p = 0 .. P-2
TBP_HD void pb_tails(const PartView &v, int64_t seg, const double *out) {
    if (v.p_start[seg] < 0) return;
    const int64_t g0 = v.seg_chunk0[seg], P = v.seg_chunk0[seg + 1] - g0;
    const int64_t q = v.p_width[seg] - 1, m = v.seg_m[seg], s0 = v.seg_start[seg];
    const double *Gf = v.Gf + v.g_off[seg];
    for (int64_t p = 0; p + 1 < P; ++p) {
        const int64_t s = p * m, e = s + m;
        double *t = v.tails + (g0 + p) * v.qmax;
        for (int64_t c = 0; c < q; ++c) {
            const int64_t j = e - q + c;
            const double base = out[s0 + j];
            t[c] = (p == 0) ? base
                            : chunk_correct(Gf + s * q, q + 1, s, j, base,
                                            v.tails + (g0 + p - 1) * v.qmax);
        }
    }
}

// launch 3 (thread = row j of segment seg): y = u + Gf t of the previous chunk
TBP_HD void pb_fwd_correct(const PartView &v, int64_t seg, int64_t j, double *out) {
    if (v.p_start[seg] < 0) return;
    const int64_t m = v.seg_m[seg], p = j / m;
    if (p == 0) return;
    const int64_t q = v.p_width[seg] - 1, s = p * m, g0 = v.seg_chunk0[seg];
    double *y = out + v.seg_start[seg] + j;
    *y = chunk_correct(v.Gf + v.g_off[seg] + s * q, q + 1, s, j, *y,
                       v.tails + (g0 + p - 1) * v.qmax);
}

// launch 4 (thread = chunk): chunk-local back solve, in place
TBP_HD void pb_bwd_local(const PartView &v, int64_t g, double *out) {
    const ChunkRef c = chunk_ref(v, g);
    if (c.cut) return;
    bwd_chunk(c.ab, c.w, c.n, c.s, c.e, out + c.s0, out + c.s0);
}

// launch 5 (thread = segment): heads h[p] = x[s_p : s_p + q], p = P-1 .. 1 (missing rows = 0)
TBP_HD void pb_heads(const PartView &v, int64_t seg, const double *out) {
    if (v.p_start[seg] < 0) return;
    const int64_t g0 = v.seg_chunk0[seg], P = v.seg_chunk0[seg + 1] - g0;
    const int64_t n = v.seg_len[seg], q = v.p_width[seg] - 1, m = v.seg_m[seg];
    const int64_t s0 = v.seg_start[seg];
    const double *Gb = v.Gb + v.g_off[seg];
    for (int64_t p = P - 1; p >= 1; --p) {
        const int64_t s = p * m, e = (s + m < n) ? s + m : n;
        double *h = v.heads + (g0 + p) * v.qmax;
        for (int64_t c = 0; c < q; ++c) {
            const int64_t j = s + c;
            if (j >= e) {
                h[c] = 0.0;
                continue;
            }
            const double base = out[s0 + j];
            h[c] = (p == P - 1) ? base
                                : chunk_correct(Gb + s * q, q + 1, s, j, base,
                                                v.heads + (g0 + p + 1) * v.qmax);
        }
    }
}

// launch 6 (thread = row): x = v + Gb h of the next chunk; flagged amplitudes -> 0
TBP_HD void pb_bwd_correct(const PartView &v, int64_t seg, int64_t j, const uint8_t *flags,
                           double *out) {
    const int64_t idx = v.seg_start[seg] + j;
    if (v.p_start[seg] >= 0) {
        const int64_t g0 = v.seg_chunk0[seg], P = v.seg_chunk0[seg + 1] - g0;
        const int64_t m = v.seg_m[seg], p = j / m, q = v.p_width[seg] - 1;
        if (p + 1 < P) {
            const int64_t s = p * m;
            out[idx] = chunk_correct(v.Gb + v.g_off[seg] + s * q, q + 1, s, j, out[idx],
                                     v.heads + (g0 + p + 1) * v.qmax);
        }
    }
    if (flags[idx] != 0) out[idx] = 0.0;
}

// Host-side tables of the partitioned solve: chunking of every segment and the response arrays of
// every chunk (once per factor).  `chunk` is the requested chunk length; a segment uses
// max(chunk, width - 1).
struct PartTables {
    std::vector<int64_t> seg_chunk0, seg_m, chunk_seg, g_off;
    std::vector<double> Gf, Gb;
    int64_t n_chunk = 0, qmax = 1;
};

inline void build_part_tables(int64_t n_seg, const int64_t *seg_len, const int64_t *p_start,
                              const int64_t *p_width, const double *factors, int64_t chunk,
                              PartTables &T) {
    T.seg_chunk0.assign(n_seg + 1, 0);
    T.seg_m.assign(n_seg, 1);
    T.g_off.assign(n_seg, 0);
    T.chunk_seg.clear();
    int64_t goff = 0;
    T.qmax = 1;
    for (int64_t s = 0; s < n_seg; ++s) {
        const int64_t n = seg_len[s];
        const bool cut = p_start[s] < 0;
        const int64_t q = cut ? 0 : p_width[s] - 1;
        int64_t m = chunk > q ? chunk : q;
        if (m < 1) m = 1;
        T.seg_m[s] = m;
        const int64_t P = n > 0 ? (n + m - 1) / m : 0;
        T.seg_chunk0[s] = (int64_t)T.chunk_seg.size();
        for (int64_t p = 0; p < P; ++p) T.chunk_seg.push_back(s);
        T.g_off[s] = goff;
        goff += n * q;
        if (q > T.qmax) T.qmax = q;
    }
    T.seg_chunk0[n_seg] = (int64_t)T.chunk_seg.size();
    T.n_chunk = (int64_t)T.chunk_seg.size();
    T.Gf.assign((size_t)(goff > 0 ? goff : 1), 0.0);
    T.Gb.assign((size_t)(goff > 0 ? goff : 1), 0.0);
    for (int64_t s = 0; s < n_seg; ++s) {
        if (p_start[s] < 0) continue;
        const int64_t n = seg_len[s], w = p_width[s], q = w - 1, m = T.seg_m[s];
        if (q == 0) continue;
        const double *ab = factors + p_start[s];
        const int64_t P = T.seg_chunk0[s + 1] - T.seg_chunk0[s];
        for (int64_t p = 0; p < P; ++p) {
            const int64_t a = p * m, e = (a + m < n) ? a + m : n;
            if (p > 0) fwd_response(ab, w, n, a, e, T.Gf.data() + T.g_off[s] + a * q);
            if (p + 1 < P) bwd_response(ab, w, n, a, e, T.Gb.data() + T.g_off[s] + a * q);
        }
    }
}

} // namespace tbp
