// tb_wcs.cuh -- detector quaternion -> flat-projection pixel (ops/pixels_wcs.py:39-662).
//
// The reference does this on the HOST, per detector and view:
//   lon, lat  = center_offset_lonlat(quats)          pointing_utils.py:16-67 -> qa_to_iso
//                                                     (_libtoast/math_qarray.cpp:694-783)
//   col, row  = around(wcs.wcs_world2pix([lon, lat] in degrees, origin 0))    astropy / WCSLIB
//   pixel     = col + row * n_col;  >= n_pix or flagged -> -1
// WCSLIB (astropy's bundled copy; a third-party dependency that is not part of /root/reference) is
// restated here from its published algorithm -- Calabretta & Greisen 2002, "Representations of
// celestial coordinates in FITS" (Paper II): celestial -> native spherical rotation with the
// Euler angles of `celset`, the projection equations of CAR, CEA, MER, SFL (cylindrical /
// pseudo-cylindrical, native reference point (0, 0)) and TAN, ZEA (zenithal, native pole at the
// reference point), the degree-argument trigonometry of wcstrig.c with its exact values at
// multiples of 90 degrees, and the linear step p = x / CDELT + CRPIX (unit PC matrix).
// PARITY UNPINNED against WCSLIB itself (astropy is not in this image); pinned against the
// projection formulas, the reference test's pixel-centre round trip (tests/ops_pointing_wcs.py:
// 45-78, 165-215) and the numpy restatement in oracle/pixels_wcs.py.
//
// Everything is host/device code so that the CPU suite runs the same functions the kernel runs.
#pragma once

#include "tb_math.cuh"

namespace tbw {

constexpr double kPi = 3.141592653589793238462643;
constexpr double kD2R = kPi / 180.0;
constexpr double kR2D = 180.0 / kPi;
constexpr double kTrigTol = 1e-10; // WCSTRIG_TOL

enum Proj { CAR = 0, CEA = 1, MER = 2, SFL = 3, TAN = 4, ZEA = 5 };

// ---- wcstrig.c -----------------------------------------------------------------------------
TB_HD void sincosd(double angle, double &s, double &c) {
    if (fmod(angle, 90.0) == 0.0) {
        int i = (int)floor(angle / 90.0 + 0.5);
        i = (i < 0 ? -i : i) % 4;
        switch (i) {
        case 0: s = 0.0; c = 1.0; return;
        case 1: s = (angle > 0.0) ? 1.0 : -1.0; c = 0.0; return;
        case 2: s = 0.0; c = -1.0; return;
        default: s = (angle > 0.0) ? -1.0 : 1.0; c = 0.0; return;
        }
    }
    s = sin(angle * kD2R);
    c = cos(angle * kD2R);
}
TB_HD double cosd(double angle) {
    double s, c;
    sincosd(angle, s, c);
    return c;
}
TB_HD double sind(double angle) {
    double s, c;
    sincosd(angle, s, c);
    return s;
}
TB_HD double tand(double angle) {
    double resid = fmod(angle, 360.0);
    if (resid == 0.0 || fabs(resid) == 180.0) return 0.0;
    if (resid == 45.0 || resid == 225.0) return 1.0;
    if (resid == -135.0 || resid == -315.0) return -1.0;
    return tan(angle * kD2R);
}
TB_HD double acosd(double v) {
    if (v >= 1.0) {
        if (v - 1.0 < kTrigTol) return 0.0;
    } else if (v == 0.0) {
        return 90.0;
    } else if (v <= -1.0) {
        if (v + 1.0 > -kTrigTol) return 180.0;
    }
    return acos(v) * kR2D;
}
TB_HD double asind(double v) {
    if (v <= -1.0) {
        if (v + 1.0 > -kTrigTol) return -90.0;
    } else if (v == 0.0) {
        return 0.0;
    } else if (v >= 1.0) {
        if (v - 1.0 < kTrigTol) return 90.0;
    }
    return asin(v) * kR2D;
}
TB_HD double atan2d(double y, double x) {
    if (y == 0.0) {
        if (x >= 0.0) return 0.0;
        if (x < 0.0) return 180.0;
    } else if (x == 0.0) {
        if (y > 0.0) return 90.0;
        if (y < 0.0) return -90.0;
    }
    return atan2(y, x) * kR2D;
}

// ---- the projection as the kernel sees it ---------------------------------------------------
struct Wcs {
    int proj;
    double euler[5];  // celset: lng_p, 90 - lat_p, phi_p, cos / sin of euler[1]
    double crpix[2], cdelt[2];
    double cea_lambda; // CEA: PV2_1
    int64_t n_col, n_pix;
    int is_azimuth;
};

// celestial (lng, lat) -> native (phi, theta), all in degrees: sphs2x
TB_HD void sph_s2x(const double *eul, double lng, double lat, double &phi, double &theta) {
    const double tol = 1.0e-5;
    if (eul[4] == 0.0) {
        if (eul[1] == 0.0) {
            double dphi = fmod(eul[2] - 180.0 - eul[0], 360.0);
            phi = fmod(lng + dphi, 360.0);
            theta = lat;
        } else {
            double dphi = fmod(eul[2] + eul[0], 360.0);
            phi = fmod(dphi - lng, 360.0);
            theta = -lat;
        }
        if (phi > 180.0) phi -= 360.0;
        else if (phi < -180.0) phi += 360.0;
        return;
    }
    const double dlng = lng - eul[0];
    double sinlng, coslng, sinlat, coslat;
    sincosd(dlng, sinlng, coslng);
    sincosd(lat, sinlat, coslat);
    const double coslat3 = coslat * eul[3], coslat4 = coslat * eul[4];
    const double sinlat3 = sinlat * eul[3], sinlat4 = sinlat * eul[4];
    double x = sinlat4 - coslat3 * coslng;
    if (fabs(x) < tol) x = -cosd(lat + eul[1]) + coslat3 * (1.0 - coslng);
    const double y = -coslat * sinlng;
    double dphi;
    if (x != 0.0 || y != 0.0) {
        dphi = atan2d(y, x);
    } else {
        dphi = (eul[1] < 90.0) ? dlng - 180.0 : -dlng;
    }
    phi = fmod(eul[2] + dphi, 360.0);
    if (phi > 180.0) phi -= 360.0;
    else if (phi < -180.0) phi += 360.0;
    if (fmod(dlng, 180.0) == 0.0) {
        theta = lat + coslng * eul[1];
        if (theta > 90.0) theta = 180.0 - theta;
        if (theta < -90.0) theta = -180.0 - theta;
    } else {
        const double z = sinlat3 + coslat4 * coslng;
        if (fabs(z) > 0.99) {
            const double a = acosd(sqrt(x * x + y * y));
            theta = (z < 0.0) ? -fabs(a) : fabs(a);
        } else {
            theta = asind(z);
        }
    }
}

// native (phi, theta) -> projection plane (x, y) in degrees (r0 = 180 / pi); false = the point
// has no image (TAN beyond the horizon, MER at the pole)
TB_HD bool prj_s2x(int proj, double lambda, double phi, double theta, double &x, double &y) {
    switch (proj) {
    case CAR:
        x = phi;
        y = theta;
        return true;
    case CEA:
        x = phi;
        y = (kR2D / lambda) * sind(theta);
        return true;
    case MER:
        x = phi;
        if (theta <= -90.0) return false;
        y = kR2D * log(tand((theta + 90.0) / 2.0));
        return true;
    case SFL:
        x = phi * cosd(theta);
        y = theta;
        return true;
    case TAN: {
        const double s = sind(theta);
        if (s <= 0.0) return false;
        const double r = kR2D * cosd(theta) / s;
        double sp, cp;
        sincosd(phi, sp, cp);
        x = r * sp;
        y = -r * cp;
        return true;
    }
    default: { // ZEA
        const double r = 2.0 * kR2D * sind((90.0 - theta) / 2.0);
        double sp, cp;
        sincosd(phi, sp, cp);
        x = r * sp;
        y = -r * cp;
        return true;
    }
    }
}

// world (degrees) -> fractional pixel, origin 0 (wcs_world2pix(..., 0))
TB_HD bool world2pix(const Wcs &w, double lng, double lat, double &col, double &row) {
    double phi, theta, x, y;
    sph_s2x(w.euler, lng, lat, phi, theta);
    if (!prj_s2x(w.proj, w.cea_lambda, phi, theta, x, y)) return false;
    col = (x / w.cdelt[0] + w.crpix[0]) - 1.0;
    row = (y / w.cdelt[1] + w.crpix[1]) - 1.0;
    return true;
}

// detector quaternion -> (lon, lat) in degrees: qa_to_iso + center_offset_lonlat without a moving
// centre (math_qarray.cpp:738-772, pointing_utils.py:35-66)
TB_HD void quat_to_lonlat_deg(const tbm::Quat &qin, int is_azimuth, double &lon, double &lat) {
    const double pi = 3.14159265358979323846, pi_2 = 1.57079632679489661923;
    const double eps = 2.220446049250313e-16;
    double norm = 0.0;
    norm += qin.x * qin.x;
    norm += qin.y * qin.y;
    norm += qin.z * qin.z;
    norm += qin.w * qin.w;
    norm = 1.0 / sqrt(norm);
    tbm::Quat q{qin.x * norm, qin.y * norm, qin.z * norm, qin.w * norm};
    double dx, dy, dz;
    tbm::rot_zaxis(q, dx, dy, dz);
    double theta, phi;
    if (fabs(fabs(dz) - 1.0) < eps) {
        phi = 0.0;
        theta = (dz >= 0.0) ? 0.0 : pi;
    } else {
        theta = pi_2 - asin(dz);
        phi = atan2(dy, dx);
    }
    double lon_rad = phi;
    const double lat_rad = 0.5 * pi - theta;
    if (is_azimuth) lon_rad = 2 * pi - lon_rad;
    if (lon_rad >= 2 * pi) lon_rad -= 2 * pi;
    if (lon_rad < 0) lon_rad += 2 * pi;
    lon = lon_rad * (180.0 / pi); // np.degrees
    lat = lat_rad * (180.0 / pi);
}

// the whole per-sample chain; -1 for samples without an image or beyond the last pixel.  (A
// negative col / row that still yields a value below n_pix is passed through unchanged, as the
// reference does: only `pixels >= n_pix` is tested, ops/pixels_wcs.py:611-616.)
TB_HD int64_t quat_to_wcs_pixel(const Wcs &w, const tbm::Quat &q) {
    double lon, lat, dcol, drow;
    quat_to_lonlat_deg(q, w.is_azimuth, lon, lat);
    if (!world2pix(w, lon, lat, dcol, drow)) return -1;
    const int64_t col = (int64_t)rint(dcol), row = (int64_t)rint(drow); // np.around: half to even
    const int64_t p = col + row * w.n_col;
    return (p >= w.n_pix) ? -1 : p;
}

} // namespace tbw
