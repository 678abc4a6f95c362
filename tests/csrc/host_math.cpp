// Test-only host build of toast_b200/csrc/tb_math.cuh (g++ -ffp-contract=off): lets the CPU
// test-suite check the product's per-sample arithmetic (exact atan2, two-tier pixel path,
// trig-free Stokes weights) against glibc and the oracle without a GPU.
#include "../../toast_b200/csrc/tb_math.cuh"
#include "../../toast_b200/csrc/tb_wcs.cuh"
#include <cstdint>

extern "C" {

void tbm_atan2_cr(int64_t n, const double *y, const double *x, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = tbm::atan2_cr(y[i], x[i]);
}

// quats [n,4] -> pixels; mode 0 = two-tier with guard_scale, 1 = exact path for every sample
int64_t tbm_quat2pix(int64_t n, const double *quats, int64_t nside, int nest, double guard_scale,
                     int64_t *pix) {
    tbm::PixCtx c = tbm::make_pix_ctx(nside, guard_scale);
    int64_t n_exact = 0;
    for (int64_t i = 0; i < n; ++i) {
        tbm::Quat q{quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]};
        double dx, dy, dz;
        tbm::rot_zaxis(q, dx, dy, dz);
        int ex = 0;
        pix[i] = nest ? tbm::vec2pix<true>(c, dx, dy, dz, &ex) : tbm::vec2pix<false>(c, dx, dy, dz, &ex);
        n_exact += ex;
    }
    return n_exact;
}

// pixel from (z, phi) directly: exercises zphi2pix with an externally supplied phi
void tbm_zphi2pix(int64_t n, const double *z, const double *phi, int64_t nside, int nest,
                  int64_t *pix, uint8_t *amb) {
    tbm::PixCtx c = tbm::make_pix_ctx(nside, 1.0);
    for (int64_t i = 0; i < n; ++i) {
        bool a = false;
        pix[i] = c.small ? (nest ? tbm::zphi2pix<true, int32_t>(c, phi[i], z[i], a) : tbm::zphi2pix<false, int32_t>(c, phi[i], z[i], a))
                         : (nest ? tbm::zphi2pix<true, int64_t>(c, phi[i], z[i], a) : tbm::zphi2pix<false, int64_t>(c, phi[i], z[i], a));
        amb[i] = a ? 1 : 0;
    }
}

// the per-pixel 3x3 eigen-inverse of k_cov_invert3 over a whole map
void tbm_cov_invert3(int64_t npix, double *cov, double *rcond, double threshold) {
    for (int64_t p = 0; p < npix; ++p) rcond[p] = tbm::cov_invert3(cov + 6 * p, threshold);
}

void tbm_stokes_iqu(int64_t n, const double *quats, double cal, double eps, double U_sign,
                    double gamma, const double *hwp, double *w) {
    double eta = (1.0 - eps) / (1.0 + eps);
    for (int64_t i = 0; i < n; ++i) {
        tbm::Quat q{quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]};
        if (hwp)
            tbm::stokes_iqu<true>(q, cal, eta, U_sign, gamma, hwp[i], w[3 * i], w[3 * i + 1], w[3 * i + 2]);
        else
            tbm::stokes_iqu<false>(q, cal, eta, U_sign, gamma, 0.0, w[3 * i], w[3 * i + 1], w[3 * i + 2]);
    }
}
}

// ---- Offset noise prior cores (toast_b200/csrc/tb_prior.cuh), looped the way the kernels do ----
#include "../../toast_b200/csrc/tb_prior.cuh"

extern "C" {

// mode 0: out += conv, flagged -> 0 (add_prior); mode 1: out = conv, flagged -> 0 (Toeplitz)
void tbp_conv_segments(int64_t n_seg, const int64_t *seg_start, const int64_t *seg_len,
                       const int64_t *f_start, const int64_t *f_len, const double *taps,
                       const double *in, const uint8_t *flags, double *out, int mode) {
    for (int64_t s = 0; s < n_seg; ++s) {
        for (int64_t i = 0; i < seg_len[s]; ++i) {
            const int64_t g = seg_start[s] + i;
            double v = 0.0;
            if (f_start[s] >= 0 && flags[g] == 0) {
                v = tbp::conv_same_at(in + seg_start[s], seg_len[s], taps + f_start[s], f_len[s], i);
                if (mode == 0) v += out[g];
            }
            out[g] = v;
        }
    }
}

void tbp_banded_segments(int64_t n_seg, const int64_t *seg_start, const int64_t *seg_len,
                         const int64_t *p_start, const int64_t *p_width, const double *factors,
                         const double *in, const uint8_t *flags, double *out) {
    for (int64_t s = 0; s < n_seg; ++s) {
        const int64_t n = seg_len[s], s0 = seg_start[s];
        if (p_start[s] < 0) {
            for (int64_t j = 0; j < n; ++j) out[s0 + j] = 0.0;
            continue;
        }
        tbp::banded_cho_solve(factors + p_start[s], p_width[s], n, in + s0, out + s0);
        for (int64_t j = 0; j < n; ++j)
            if (flags[s0 + j] != 0) out[s0 + j] = 0.0;
    }
}
}

// ---- partitioned banded solve (tb_prior.cuh), looped the way the device kernels would ----------
#include <vector>

extern "C" {

// one segment: factor ab [w, n], chunk size m >= w - 1, right-hand side b -> x
void tbp_banded_partitioned(const double *ab, int64_t w, int64_t n, int64_t m, const double *b,
                            double *x) {
    const int64_t q = w - 1;
    const int64_t P = (n + m - 1) / m;
    std::vector<double> Gf((size_t)(n * q), 0.0), Gb((size_t)(n * q), 0.0);
    // set-up (once per factor): responses of every chunk
    for (int64_t p = 0; p < P; ++p) {
        const int64_t s = p * m, e = (s + m < n) ? s + m : n;
        if (p > 0) tbp::fwd_response(ab, w, n, s, e, Gf.data() + s * q);
        if (p + 1 < P) tbp::bwd_response(ab, w, n, s, e, Gb.data() + s * q);
    }
    std::vector<double> t((size_t)(P * (q > 0 ? q : 1)), 0.0);
    // K1: chunk-local forward solves (parallel over chunks)
    for (int64_t p = 0; p < P; ++p) {
        const int64_t s = p * m, e = (s + m < n) ? s + m : n;
        tbp::fwd_chunk(ab, w, n, s, e, b, x);
    }
    // K2: tails, sequential over chunks; t[p] = y[e_p - q : e_p]
    for (int64_t p = 0; p + 1 < P; ++p) {
        const int64_t s = p * m, e = s + m;
        for (int64_t c = 0; c < q; ++c) {
            const int64_t j = e - q + c;
            t[p * q + c] = (p == 0) ? x[j]
                                    : tbp::chunk_correct(Gf.data() + s * q, w, s, j, x[j],
                                                         t.data() + (p - 1) * q);
        }
    }
    // K3: correction (parallel over rows)
    for (int64_t p = 1; p < P; ++p) {
        const int64_t s = p * m, e = (s + m < n) ? s + m : n;
        for (int64_t j = s; j < e; ++j)
            x[j] = tbp::chunk_correct(Gf.data() + s * q, w, s, j, x[j], t.data() + (p - 1) * q);
    }
    // K4: chunk-local back solves
    for (int64_t p = 0; p < P; ++p) {
        const int64_t s = p * m, e = (s + m < n) ? s + m : n;
        tbp::bwd_chunk(ab, w, n, s, e, x, x);
    }
    // K5: heads, sequential from the last chunk; h[p] = x[s_p : s_p + q] (rows >= n are zero)
    std::vector<double> h((size_t)(P * (q > 0 ? q : 1)), 0.0);
    for (int64_t p = P - 1; p >= 1; --p) {
        const int64_t s = p * m, e = (s + m < n) ? s + m : n;
        for (int64_t c = 0; c < q; ++c) {
            const int64_t j = s + c;
            if (j >= e) {
                h[p * q + c] = 0.0;
                continue;
            }
            h[p * q + c] = (p == P - 1) ? x[j]
                                        : tbp::chunk_correct(Gb.data() + s * q, w, s, j, x[j],
                                                             h.data() + (p + 1) * q);
        }
    }
    // K6: correction
    for (int64_t p = 0; p + 1 < P; ++p) {
        const int64_t s = p * m, e = s + m;
        for (int64_t j = s; j < e; ++j)
            x[j] = tbp::chunk_correct(Gb.data() + s * q, w, s, j, x[j], h.data() + (p + 1) * q);
    }
}
}

// ---- the partitioned solve exactly as tb_prior.cu launches it: tables + six per-thread loops ----
extern "C" {
void tbp_banded_partitioned_segments(int64_t n_seg, const int64_t *seg_start,
                                     const int64_t *seg_len, const int64_t *p_start,
                                     const int64_t *p_width, const double *factors, int64_t chunk,
                                     const double *in, const uint8_t *flags, double *out) {
    tbp::PartTables T;
    tbp::build_part_tables(n_seg, seg_len, p_start, p_width, factors, chunk, T);
    std::vector<double> tails((size_t)(T.n_chunk * T.qmax + 1), 0.0);
    std::vector<double> heads((size_t)(T.n_chunk * T.qmax + 1), 0.0);
    tbp::PartView v;
    v.n_seg = n_seg;
    v.n_chunk = T.n_chunk;
    v.qmax = T.qmax;
    v.seg_start = seg_start;
    v.seg_len = seg_len;
    v.p_start = p_start;
    v.p_width = p_width;
    v.factors = factors;
    v.seg_chunk0 = T.seg_chunk0.data();
    v.seg_m = T.seg_m.data();
    v.chunk_seg = T.chunk_seg.data();
    v.g_off = T.g_off.data();
    v.Gf = T.Gf.data();
    v.Gb = T.Gb.data();
    v.tails = tails.data();
    v.heads = heads.data();
    for (int64_t g = 0; g < T.n_chunk; ++g) tbp::pb_fwd_local(v, g, in, out);
    for (int64_t s = 0; s < n_seg; ++s) tbp::pb_tails(v, s, out);
    for (int64_t s = 0; s < n_seg; ++s)
        for (int64_t j = 0; j < seg_len[s]; ++j) tbp::pb_fwd_correct(v, s, j, out);
    for (int64_t g = 0; g < T.n_chunk; ++g) tbp::pb_bwd_local(v, g, out);
    for (int64_t s = 0; s < n_seg; ++s) tbp::pb_heads(v, s, out);
    for (int64_t s = 0; s < n_seg; ++s)
        for (int64_t j = 0; j < seg_len[s]; ++j) tbp::pb_bwd_correct(v, s, j, flags, out);
}

// tb_wcs.cuh on the host: quats [n,4] -> flat-projection pixels (and the fractional coordinates)
void tbw_quat2pix(int64_t n, const double *quats, int proj, int is_azimuth, const double *euler,
                  const double *crpix, const double *cdelt, double cea_lambda, int64_t n_col,
                  int64_t n_row, int64_t *pix, double *dcol, double *drow) {
    tbw::Wcs w;
    w.proj = proj;
    w.is_azimuth = is_azimuth;
    for (int k = 0; k < 5; ++k) w.euler[k] = euler[k];
    for (int k = 0; k < 2; ++k) {
        w.crpix[k] = crpix[k];
        w.cdelt[k] = cdelt[k];
    }
    w.cea_lambda = cea_lambda;
    w.n_col = n_col;
    w.n_pix = n_col * n_row;
    for (int64_t i = 0; i < n; ++i) {
        tbm::Quat q{quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]};
        pix[i] = tbw::quat_to_wcs_pixel(w, q);
        double lon, lat;
        tbw::quat_to_lonlat_deg(q, is_azimuth, lon, lat);
        if (!tbw::world2pix(w, lon, lat, dcol[i], drow[i])) dcol[i] = drow[i] = 0.0;
    }
}
}
