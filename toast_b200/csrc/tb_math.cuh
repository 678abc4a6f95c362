// tb_math.cuh -- per-sample arithmetic of the TOAST pointing chain, written for sm_100a.
//
// Compiled with -fmad=false (device) / -ffp-contract=off (host test build) so that every
// a*b+c below rounds twice exactly like the reference's baseline-x86-64 build; the only FMAs
// are the explicit fma() calls inside the double-double helpers.
//
// Reference behaviour being reproduced (file:line relative to
// /root/reference/src/toast/_libtoast):
//   quaternion product                ops_pointing_detector.cpp:21-31
//   rotation of z^ / x^ by a quat     ops_pixels_healpix.cpp:50-76, ops_stokes_weights.cpp:21-49
//   vec -> (z, phi, region)           ops_pixels_healpix.cpp:104-121
//   (z, phi) -> NEST / RING pixel     ops_pixels_healpix.cpp:123-208, :210-274
//   detector angle and IQU weights    ops_stokes_weights.cpp:50-140
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TB_HD __host__ __device__ __forceinline__
#define TB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define TB_HD inline
#define TB_HD_NOINLINE __attribute__((noinline))
#endif

namespace tbm {

// ------------------------------------------------------------------------------------------
// quaternions (scalar last)
// ------------------------------------------------------------------------------------------
struct Quat {
    double x, y, z, w;
};

// r = p (x) q
TB_HD Quat qmul(const Quat &p, const Quat &q) {
    Quat r;
    r.x = p.x * q.w + p.y * q.z - p.z * q.y + p.w * q.x;
    r.y = -p.x * q.z + p.y * q.w + p.z * q.x + p.w * q.y;
    r.z = p.x * q.y - p.y * q.x + p.z * q.w + p.w * q.z;
    r.w = -p.x * q.x - p.y * q.y - p.z * q.z + p.w * q.w;
    return r;
}

// Line of sight R(q) z^ .  The reference runs the general rotation formula on (0,0,1); the
// products with 0.0 are signed zeros, adding them is exact, and its trailing "+ v_in[i]"
// (+0.0 here) turns any -0 into +0 -- so this reduced form is bit-identical.
TB_HD void rot_zaxis(const Quat &q, double &dx, double &dy, double &dz) {
    double xw = q.w * q.x, yw = q.w * q.y;
    double x2 = -q.x * q.x, y2 = -q.y * q.y;
    double xz = q.x * q.z, yz = q.y * q.z;
    dx = 2 * (yw + xz) + 0.0;
    dy = 2 * (yz - xw) + 0.0;
    dz = 2 * (x2 + y2) + 1.0;
}

// Polarisation-sensitive direction R(q) x^ .
TB_HD void rot_xaxis(const Quat &q, double &ox, double &oy, double &oz) {
    double yw = q.w * q.y, zw = q.w * q.z;
    double xy = q.x * q.y, xz = q.x * q.z;
    double y2 = -q.y * q.y, z2 = -q.z * q.z;
    ox = 2 * (y2 + z2) + 1.0;
    oy = 2 * (zw + xy) + 0.0;
    oz = 2 * (xz - yw) + 0.0;
}

// ------------------------------------------------------------------------------------------
// double-double arithmetic and a (practically) correctly rounded atan2
//
// CUDA's atan2 is accurate to 2 ulp, glibc's to < 1 ulp (0.52 ulp bound after the 2.35
// slow-path removal), so the two can disagree in the last bit.  That only matters for the
// pixel NUMBER when a pre-truncation value lands within a few ulp of an integer.  Those
// samples (about 1e-10 of all samples at nside 2048) are re-evaluated with this routine,
// whose result equals the correctly rounded value and therefore glibc's except when the true
// angle lies within ~2^-60 relative of a rounding midpoint.
// ------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};

TB_HD dd two_sum(double a, double b) {
    double s = a + b;
    double bb = s - a;
    double e = (a - (s - bb)) + (b - bb);
    return dd{s, e};
}
TB_HD dd quick_two_sum(double a, double b) {
    double s = a + b;
    double e = b - (s - a);
    return dd{s, e};
}
TB_HD dd two_prod(double a, double b) {
    double p = a * b;
    double e = fma(a, b, -p);
    return dd{p, e};
}
TB_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
TB_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
TB_HD dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
TB_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo += a.lo;
    return quick_two_sum(s.hi, s.lo);
}
TB_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}
TB_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}
TB_HD dd dd_div(dd a, dd b) {
    double q1 = a.hi / b.hi;
    dd r = dd_sub(a, dd_mul_d(b, q1));
    double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    double q3 = r.hi / b.hi;
    dd q = quick_two_sum(q1, q2);
    return dd_add_d(q, q3);
}
TB_HD dd dd_sqrt(dd a) {
    // Karp & Markstein: sqrt(a) ~= a*x + (a - (a*x)^2) * x / 2 with x = 1/sqrt(a.hi)
    if (a.hi <= 0.0) return dd{0.0, 0.0};
    double x = 1.0 / sqrt(a.hi);
    double ax = a.hi * x;
    dd t = dd_sub(a, two_prod(ax, ax));
    return two_sum(ax, t.hi * (x * 0.5));
}

// atan of a dd argument in [0, 1]: three angle halvings, then the alternating series.
TB_HD dd dd_atan_unit(dd q) {
    const dd one = dd{1.0, 0.0};
    dd t = q;
    for (int i = 0; i < 3; ++i) {
        dd den = dd_add(one, dd_sqrt(dd_add(one, dd_mul(t, t))));
        t = dd_div(t, den);
    }
    // |t| <= tan(pi/32) ~ 0.0985 : 18 terms reach 2^-113
    dd t2 = dd_mul(t, t);
    const int K = 18;
    dd s = dd_div(one, dd{(double)(2 * K + 1), 0.0});
    for (int k = K - 1; k >= 0; --k) {
        dd c = dd_div(one, dd{(double)(2 * k + 1), 0.0});
        s = dd_sub(c, dd_mul(t2, s));
    }
    dd a = dd_mul(t, s);
    return dd{a.hi * 8.0, a.lo * 8.0};
}

TB_HD double atan2_cr(double y, double x) {
    const dd pi = dd{3.141592653589793116e+00, 1.224646799147353207e-16};
    const dd pio2 = dd{1.570796326794896558e+00, 6.123233995736766036e-17};
    double ay = fabs(y), ax = fabs(x);
    if (ay == 0.0 && ax == 0.0) return signbit(x) ? copysign(pi.hi, y) : y;
    dd a;
    if (ay <= ax) {
        a = dd_atan_unit(dd_div(dd{ay, 0.0}, dd{ax, 0.0}));
    } else {
        a = dd_sub(pio2, dd_atan_unit(dd_div(dd{ax, 0.0}, dd{ay, 0.0})));
    }
    if (signbit(x)) a = dd_sub(pi, a);
    double r = a.hi + a.lo;
    return signbit(y) ? -r : r;
}

// ------------------------------------------------------------------------------------------
// HEALPix
// ------------------------------------------------------------------------------------------
struct PixCtx {
    int64_t nside;
    int64_t nm1;       // nside - 1
    int64_t fournside; // 4 nside
    int64_t ncap;      // 2 (nside^2 - nside)
    int64_t npix;      // 12 nside^2
    int factor;        // log2(nside)
    double dnside;     // (double) nside
    double halfnside;  // 0.5 nside
    double tqnside;    // 0.75 nside
    double guard;      // half-width of the "ambiguous" band around integers, units of jp/jm
    double guard_tt;   // same, in units of tt
    bool small;        // nside <= 8192: 32-bit integer arithmetic is exact
};

TB_HD PixCtx make_pix_ctx(int64_t nside, double guard_scale) {
    PixCtx c;
    c.nside = nside;
    c.nm1 = nside - 1;
    c.fournside = 4 * nside;
    c.ncap = 2 * (nside * nside - nside);
    c.npix = 12 * nside * nside;
    int f = 0;
    while (((int64_t)1 << f) < nside) ++f;
    c.factor = f;
    c.dnside = (double)nside;
    c.halfnside = 0.5 * c.dnside;
    c.tqnside = 0.75 * c.dnside;
    // |delta tt| <= 2 ulp(phi) * 2/pi + a few roundings < 2e-15; use 4x that.
    c.guard_tt = 8.0e-15 * guard_scale;
    c.guard = c.guard_tt * c.dnside;
    c.small = nside <= 8192;
    return c;
}

// Spread the low 32 bits of v onto the even bit positions (the reference's 8-bit utab
// lookups, ops_pixels_healpix.cpp:20-27,78-85, done with masks instead of a table).
TB_HD uint64_t spread_bits(uint64_t v) {
    uint64_t x = v & 0xffffffffull;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}
TB_HD uint32_t spread_bits16(uint32_t v) { // 16 bits -> 32 bits
    uint32_t x = v & 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
TB_HD int64_t xy2pix(int64_t x, int64_t y) {
    return (int64_t)(spread_bits((uint64_t)x) | (spread_bits((uint64_t)y) << 1));
}
TB_HD int32_t xy2pix(int32_t x, int32_t y) {
    return (int32_t)(spread_bits16((uint32_t)x) | (spread_bits16((uint32_t)y) << 1));
}

TB_HD bool near_integer(double v, double tol) { return fabs(v - rint(v)) <= tol; }

// phi -> tt in [0, 4)  (hpix_fmod + snap + quadrant shift, ops_pixels_healpix.cpp:44-48,132-139)
// EXACT = false (first tier only): phi / twopi is formed as phi * (1 / twopi) -- at most one ulp
// away from the IEEE quotient, i.e. one more ulp of phi inside a guard band that is sized for
// eight (make_pix_ctx); a sample this could move is flagged and redone with EXACT = true.
template <bool EXACT>
TB_HD double phi_to_tt(double phi, bool &ambiguous) {
    const double eps = 2.220446049250313e-16;
    const double tol = 10.0 * eps;
    const double twopi = 2 * 3.14159265358979323846;
    const double inv_twopi = 0.15915494309189534561;
    const double two_over_pi = 0.63661977236758134308;
    double div = EXACT ? phi / twopi : phi * inv_twopi;
    // |phi| <= pi so (int64)div == 0; the subtraction of 0.0 is kept for the sign of zero
    double phi_mod = twopi * (div - (double)((int64_t)div));
    double aphi = fabs(phi_mod);
    if (fabs(aphi - tol) <= tol * 1.0e-13) ambiguous = true;
    if ((phi_mod < tol) && (phi_mod > -tol)) phi_mod = 0.0;
    return (phi_mod >= 0.0) ? phi_mod * two_over_pi : phi_mod * two_over_pi + 4.0;
}

// (z, phi) -> pixel.  Sets `ambiguous` when a 2-ulp change of phi could alter the result.
// I is the integer type of the intermediate ring / face arithmetic: int32_t is exact for
// nside <= 8192 (12 nside^2 < 2^31) and costs a third of the emulated 64-bit integer ops;
// both instantiations perform the same arithmetic as the reference's int64 code.
template <bool NEST, typename I, bool EXACT = true>
TB_HD int64_t zphi2pix(const PixCtx &c, double phi, double z, bool &ambiguous) {
    const double twothirds = 0.66666666666666666667;
    const I nside = (I)c.nside, nm1 = (I)c.nm1, fournside = (I)c.fournside;
    double za = fabs(z);
    double tt = phi_to_tt<EXACT>(phi, ambiguous);
    if (za <= twothirds) {
        double t1 = c.halfnside + c.dnside * tt;
        double t2 = c.tqnside * z;
        double vp = t1 - t2;
        double vm = t1 + t2;
        if (near_integer(vp, c.guard) || near_integer(vm, c.guard)) ambiguous = true;
        I jp = (I)vp;
        I jm = (I)vm;
        if (NEST) {
            I ifp = jp >> c.factor;
            I ifm = jm >> c.factor;
            I face;
            if (ifp == ifm) {
                face = (ifp == 4) ? (I)4 : ifp + 4;
            } else if (ifp < ifm) {
                face = ifp;
            } else {
                face = ifm + 8;
            }
            I x = jm & nm1;
            I y = nm1 - (jp & nm1);
            return (int64_t)(xy2pix(x, y) + (face << (2 * c.factor)));
        } else {
            I ir = (nside + 1) + jp - jm;
            I kshift = 1 - (ir & 1);
            I ip = (jp + jm - nside + kshift + 1) >> 1;
            ip = (ip >= 0) ? (ip & (fournside - 1)) : (ip % fournside); // fournside = 2^k
            return (int64_t)((I)c.ncap + ((ir - 1) * fournside + ip));
        }
    } else {
        // (za can exceed 1 by an ulp at the exact south pole -- dz = 1 - 2 (x^2 + y^2) with the sum
        // rounded up: the root is then NaN and so are vp / vm.  Their integer conversions give
        // INT_MIN on x86 (the reference) and 0 on the device; both end at the same pixel -- NEST
        // only uses the low bits, which are zero either way, and RING wraps to ir = 1 -- which
        // tests/test_host_math.py pins with the quaternion (s, s, 0, 0).)
        double rtz = sqrt(3.0 * (1.0 - za));
        if (near_integer(tt, c.guard_tt)) ambiguous = true;
        double t1 = c.dnside * rtz;
        if (NEST) {
            I ntt = (I)tt;
            double tp = tt - (double)ntt;
            double vp = tp * t1;
            double vm = (1.0 - tp) * t1;
            if (near_integer(vp, c.guard) || near_integer(vm, c.guard)) ambiguous = true;
            I jp = (I)vp;
            I jm = (I)vm;
            if (jp >= nside) jp = nm1;
            if (jm >= nside) jm = nm1;
            I face, x, y;
            if (z >= 0) {
                face = ntt;
                x = nm1 - jm;
                y = nm1 - jp;
            } else {
                face = ntt + 8;
                x = jp;
                y = jm;
            }
            return (int64_t)(xy2pix(x, y) + (face << (2 * c.factor)));
        } else {
            double tp = tt - floor(tt);
            double vp = tp * t1;
            double vm = (1.0 - tp) * t1;
            if (near_integer(vp, c.guard) || near_integer(vm, c.guard)) ambiguous = true;
            I jp = (I)vp;
            I jm = (I)vm;
            I ir = jp + jm + 1;
            double vi = tt * (double)ir;
            if (near_integer(vi, c.guard)) ambiguous = true;
            I ip = (I)vi;
            I longpart = (I)(ip / (4 * ir));
            ip -= longpart;
            return (z > 0.0) ? (int64_t)(2 * ir * (ir - 1) + ip)
                             : (int64_t)((I)c.npix - 2 * ir * (ir + 1) + ip);
        }
    }
}

// Out-of-line tiers: (i) the exact double-double atan2 path, taken by ~1e-10 of the samples,
// and (ii) the 64-bit-integer path for nside > 8192.  Inlining either into the unrolled hot loop
// multiplies its code size and thrashes the instruction cache.  They take scalars by value and
// rebuild the context: a reference to the kernel-parameter struct would force a local-memory
// copy of it (and a stack frame) into every caller.
template <bool NEST>
TB_HD_NOINLINE int64_t vec2pix_exact(int64_t nside, double dx, double dy, double dz) {
    PixCtx c = make_pix_ctx(nside, 1.0);
    bool dummy = false;
    double phi = atan2_cr(dy, dx);
    return c.small ? zphi2pix<NEST, int32_t>(c, phi, dz, dummy)
                   : zphi2pix<NEST, int64_t>(c, phi, dz, dummy);
}
// returns the pixel with the "ambiguous" flag in bit 62
template <bool NEST>
TB_HD_NOINLINE int64_t zphi2pix_wide(int64_t nside, double guard_tt, double phi, double z) {
    PixCtx c = make_pix_ctx(nside, 1.0);
    c.guard_tt = guard_tt;
    c.guard = guard_tt * c.dnside;
    bool amb = false;
    int64_t p = zphi2pix<NEST, int64_t>(c, phi, z, amb);
    return p | (amb ? ((int64_t)1 << 62) : 0);
}

// Direction vector -> pixel, two-tier: library atan2 first, exact atan2 only if ambiguous.
// `took_exact` (may be null) is set when the sample took the exact path.
template <bool NEST>
TB_HD int64_t vec2pix(const PixCtx &c, double dx, double dy, double dz, int *took_exact) {
    bool amb = false;
    double phi = atan2(dy, dx);
    int64_t p;
    if (c.small) {
        p = zphi2pix<NEST, int32_t, false>(c, phi, dz, amb);
    } else {
        p = zphi2pix_wide<NEST>(c.nside, c.guard_tt, phi, dz);
        amb = (p >> 62) & 1;
        p &= ~((int64_t)1 << 62);
    }
    if (amb) {
        p = vec2pix_exact<NEST>(c.nside, dx, dy, dz);
        if (took_exact) *took_exact = 1;
    }
    return p;
}

// ------------------------------------------------------------------------------------------
// Stokes weights
//
// The reference forms ang_xy = atan2(vd.y, vd.x), vm = (vd.z cos, vd.z sin, -sqrt(1-vd.z^2)),
// alpha = atan2(vd.(vm x vo), vm.vo) and then cos/sin(2 alpha).  (cos, sin)(ang_xy) is
// (vd.x, vd.y)/hypot and (cos, sin)(2 alpha) is a rational function of (alpha_x, alpha_y), so
// no transcendental call is needed: the result differs from the libm chain by ~1e-15 absolute
// (tests/test_host_math.py), five orders inside the 1e-10 parity bar.
// ------------------------------------------------------------------------------------------
TB_HD void detector_cs2alpha(double dx, double dy, double dz, double ox, double oy, double oz,
                             double &c2a, double &s2a) {
    double r2 = dx * dx + dy * dy;
    double cx = 1.0, sx = 0.0; // atan2(0, 0) = 0 in the reference
    double vm_z;
    if (r2 > 0.0) {
#ifdef __CUDA_ARCH__
        double rinv = rsqrt(r2); // 1 ulp; no division
        // sqrt(1 - dz^2) = |(dx, dy)| = r2 * rsqrt(r2) for the unit vector vd: no second square
        // root (and no cancellation near the poles, where the reference's 1 - dz^2 loses digits;
        // the weights agree with it to ~1e-15 absolute, the parity bar is 1e-10)
        vm_z = -(r2 * rinv);
#else
        double rinv = 1.0 / sqrt(r2);
        vm_z = -sqrt(1.0 - dz * dz);
#endif
        cx = dx * rinv;
        sx = dy * rinv;
    } else {
        vm_z = -sqrt(1.0 - dz * dz);
    }
    double vm_x = dz * cx;
    double vm_y = dz * sx;
    double ay = (dx * (vm_y * oz - vm_z * oy) - dy * (vm_x * oz - vm_z * ox) +
                 dz * (vm_x * oy - vm_y * ox));
    double ax = (vm_x * ox + vm_y * oy + vm_z * oz);
    // (ax, ay) = |vm x ...| (cos alpha, sin alpha) with vm, vo unit vectors orthogonal to vd, so
    // n2 = ax^2 + ay^2 = 1 + O(1e-15): one Newton step 1/n2 ~= 2 - n2 is exact to O(1e-30).
    double n2 = ax * ax + ay * ay;
    c2a = 1.0;
    s2a = 0.0; // atan2(0, 0) = 0
    if (n2 > 0.0) {
        double inv = (fabs(n2 - 1.0) < 1.0e-6) ? (2.0 - n2) : (1.0 / n2);
        c2a = (ax * ax - ay * ay) * inv;
        s2a = (2.0 * ax * ay) * inv;
    }
}

// IQU weights of one sample.  `eta_cal` = eta*cal, hwp4 = 4*(gamma - hwp[s]) (ignored if !HWP).
template <bool HWP>
TB_HD void stokes_iqu(const Quat &q, double cal, double eta, double U_sign, double gamma,
                      double hwp, double &w0, double &w1, double &w2) {
    double dx, dy, dz, ox, oy, oz;
    rot_zaxis(q, dx, dy, dz);
    rot_xaxis(q, ox, oy, oz);
    double c2a, s2a;
    detector_cs2alpha(dx, dy, dz, ox, oy, oz, c2a, s2a);
    w0 = cal;
    if (!HWP) {
        w1 = c2a * eta * cal;
        w2 = s2a * eta * cal * U_sign;
    } else {
        // ang = 2 (2 (gamma - hwp) - alpha) = b - 2 alpha,  b = 4 (gamma - hwp)
        double b = 2.0 * (2.0 * (gamma - hwp));
        double sb, cb;
        sincos(b, &sb, &cb);
        double cang = cb * c2a + sb * s2a;
        double sang = sb * c2a - cb * s2a;
        w1 = cang * eta * cal;
        w2 = -sang * eta * cal * U_sign;
    }
}

// ------------------------------------------------------------------------------------------
// Pixel covariance: symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations, then the
// reference's inverse = V diag(1/lambda) V^T and rcond = lambda_min / lambda_max
// (toast_map_cov.cpp:246-396 does the same through LAPACK dsyev + dgemm).  `m` holds the upper
// triangle row-major (6 values) and is replaced by the inverse, or by zeros when rcond is below
// the threshold; returns the rcond that is stored (0 for a rejected pixel).
// ------------------------------------------------------------------------------------------
TB_HD double cov_invert3(double *m, double threshold) {
    double a[3][3] = {{m[0], m[1], m[2]}, {m[1], m[3], m[4]}, {m[2], m[4], m[5]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double offn = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        double dn = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (offn <= 1.0e-300 || offn <= 1.0e-18 * dn) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int i = (pq == 2) ? 1 : 0;
            const int j = (pq == 0) ? 1 : 2;
            double apq = a[i][j];
            if (apq == 0.0) continue;
            double theta = (a[j][j] - a[i][i]) / (2.0 * apq);
            double t = ((theta >= 0.0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0);
            double sn = t * c;
#pragma unroll
            for (int k = 0; k < 3; ++k) { // A <- A J
                double aki = a[k][i], akj = a[k][j];
                a[k][i] = c * aki - sn * akj;
                a[k][j] = sn * aki + c * akj;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { // A <- J^T A
                double aik = a[i][k], ajk = a[j][k];
                a[i][k] = c * aik - sn * ajk;
                a[j][k] = sn * aik + c * ajk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double vki = v[k][i], vkj = v[k][j];
                v[k][i] = c * vki - sn * vkj;
                v[k][j] = sn * vki + c * vkj;
            }
        }
    }
    double e0 = a[0][0], e1 = a[1][1], e2 = a[2][2];
    double emin = fmin(e0, fmin(e1, e2));
    double emax = fmax(e0, fmax(e1, e2));
    double rc = (emax > 0.0) ? (emin / emax) : 0.0;
    bool ok = rc >= threshold;
    if (ok) {
        double i0 = 1.0 / e0, i1 = 1.0 / e1, i2 = 1.0 / e2;
        int o = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c2 = r; c2 < 3; ++c2)
                m[o++] = v[r][0] * i0 * v[c2][0] + v[r][1] * i1 * v[c2][1] + v[r][2] * i2 * v[c2][2];
    } else {
#pragma unroll
        for (int o = 0; o < 6; ++o) m[o] = 0.0;
    }
    return ok ? rc : 0.0;
}

} // namespace tbm
