// Test-only host build of toast_b200/csrc/tb_math.cuh (g++ -ffp-contract=off): lets the CPU
// test-suite check the product's per-sample arithmetic (exact atan2, two-tier pixel path,
// trig-free Stokes weights) against glibc and the oracle without a GPU.
#include "../../toast_b200/csrc/tb_math.cuh"
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
