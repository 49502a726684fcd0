// grrt_device.cuh -- device-side math of the geodesic path for sm_100a.
//
// One thread owns one ray; everything here is FP64 scalar code meant to stay in registers.
// The translation unit is compiled with -fmad=false: the reference (Rust) never contracts
// a*b+c, and the Kerr-Schild right-hand side differentiates the metric by central
// differences with h = 1e-10 (reference src/geometry/kerr.rs:162-186), which turns any
// last-bit difference in the metric into ~1e-6 relative noise in the acceleration.  Keeping
// the reference's operation order and IEEE rounding makes that noise reproduce bit for bit.
// Where the order of a sum is fixed by nalgebra (matrix products accumulate k ascending,
// the 8-component error norm pairs (0,4)(1,5)(2,6)(3,7)) it is spelled out.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "grrt_types.h"

// The same sources are compiled twice (see grrt_kern_exact.cu / grrt_kern_fast.cu): once without FMA contraction
// (bit-parity with the reference) and once with it.  GRRT_NS keeps the two sets of symbols apart.
#ifndef GRRT_NS
#define GRRT_NS grrt_exact
#endif

namespace GRRT_NS {
using namespace grrt;

#define GD __device__ __forceinline__
#define GDN __device__ __noinline__

constexpr double kPi = 3.14159265358979323846;

struct Cie {
    double x, y, z, a;
};
struct V3 {
    double x, y, z;
};

GD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
GD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
GD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
GD V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
GD V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
GD V3 operator/(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
GD double dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
GD V3 cross3(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GD double norm3(V3 a) { return sqrt(dot3(a, a)); }

// sin and cos together, ~1 ulp.  Cody-Waite reduction by pi/2 with FMA (two terms suffice for |x| < 1e5; larger
// arguments go to the library routine), fdlibm's minimax polynomials on [-pi/4, pi/4].  The coefficients live in
// the constant bank: CUDA's sincos() materialises its 14 double constants with 28 UMOV instructions per call, which
// was 18 % of the instructions of an RKF45 stage (profiles/r1_c2_trace_kernel_v3.json).  Neither routine is
// bit-identical to glibc's sin/cos; the difference (last ulp) is part of the measured parity budget.
__constant__ double c_trig[16] = {
    6.36619772367581382433e-01,                                   // 0: 2/pi
    1.57079632679489655800e+00, 6.12323399573676603587e-17,       // 1,2: pi/2 = hi + lo
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,   // 3..8: S1..S6
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,    // 9..14: C1..C6
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
    0.0};
// large or non-finite arguments: the library's Payne-Hanek path, out of line (never taken by a healthy geodesic)
__device__ __noinline__ void sincos_large(double x, double* sn, double* cs) { sincos(x, sn, cs); }
GD void sincos_pi2(double x, double* sn, double* cs) {
    if (!(fabs(x) < 1.0e5)) {
        sincos_large(x, sn, cs);
        return;
    }
    const double kd = rint(x * c_trig[0]);
    const int k = __double2int_rn(kd);
    double r = fma(-kd, c_trig[1], x);
    r = fma(-kd, c_trig[2], r);
    const double z = r * r;
    double ps = fma(z, c_trig[8], c_trig[7]);
    ps = fma(z, ps, c_trig[6]);
    ps = fma(z, ps, c_trig[5]);
    ps = fma(z, ps, c_trig[4]);
    ps = fma(z, ps, c_trig[3]);
    const double s = fma(z * r, ps, r);
    double pc = fma(z, c_trig[14], c_trig[13]);
    pc = fma(z, pc, c_trig[12]);
    pc = fma(z, pc, c_trig[11]);
    pc = fma(z, pc, c_trig[10]);
    pc = fma(z, pc, c_trig[9]);
    const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double a = (k & 1) ? c : s;    // sin(x) before the sign
    const double b = (k & 1) ? s : c;    // cos(x) before the sign
    *sn = (k & 2) ? -a : a;
    *cs = ((k + 1) & 2) ? -b : b;
}

// libm routines used by the event code (shading, chart conversions, the ray-march) as calls: CUDA inlines pow (~250
// instructions), exp, atan2, log10 and the Payne-Hanek sincos at every call site, which made the ray-march loop 4200
// instructions (67 KB) -- far beyond the instruction caches, and it is streamed once per block of 32 samples.  One shared
// body per routine keeps that code resident.  Same routines, same results.
#ifndef GRRT_NO_LIBM_CALLS
__device__ __noinline__ double m_pow(double a, double b) { return pow(a, b); }
__device__ __noinline__ double m_exp(double a) { return exp(a); }
__device__ __noinline__ double m_atan2(double a, double b) { return atan2(a, b); }
__device__ __noinline__ double m_log10(double a) { return log10(a); }
__device__ __noinline__ double m_acos(double a) { return acos(a); }
__device__ __noinline__ void m_sincos(double x, double* sn, double* cs) { sincos(x, sn, cs); }
__device__ __noinline__ double m_sin(double a) { return sin(a); }
__device__ __noinline__ double m_cos(double a) { return cos(a); }
#else
#define m_sin sin
#define m_cos cos
#define m_pow pow
#define m_exp exp
#define m_atan2 atan2
#define m_log10 log10
#define m_acos acos
#define m_sincos sincos
#endif

// the same routine as a call, for rare paths inside straight-line hot code
__device__ __noinline__ void sincos_cold(double x, double* sn, double* cs) { sincos_pi2(x, sn, cs); }

// Rust semantics
GD double rclamp(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
GD bool in01(double t) { return t >= 0.0 && t <= 1.0; }
// all four finite; one exponent test on the OR-free form: a value is finite iff |v| <= DBL_MAX (false for NaN).
// Evaluated without short-circuit branches -- the stop tests run once per accepted step.
GD bool finite4(const double* v) {
    const double m = 1.7976931348623157e308;
    return (fabs(v[0]) <= m) & (fabs(v[1]) <= m) & (fabs(v[2]) <= m) & (fabs(v[3]) <= m);
}
GD double rem_euclid(double a, double b) {
    double r = fmod(a, b);
    return r < 0.0 ? r + fabs(b) : r;
}
GD uint32_t sat_u32(double x) {
    if (!(x > 0.0)) return 0u;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}
GD unsigned long long sat_u64(double x) {
    if (!(x > 0.0)) return 0ull;
    if (x >= 18446744073709551615.0) return 0xffffffffffffffffull;
    return (unsigned long long)x;
}

#ifdef GRRT_FAST_MATH
// 1/x for the FMA build: hardware seed (rel. error 2^-23) + two Newton steps; ~1 ulp, no branches, 5 instructions
// instead of the ~14 (plus a slow-path call) of an IEEE division.  Only used where the divisor is a well-scaled
// positive/finite quantity of the right-hand sides (r, Delta, sin(theta), 1 - r_s/r).
GD double frcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}
#endif

// ---- Kerr-Schild metric pieces (kerr.rs:31-110) -------------------------------------------
GD double ks_r_sqr(double a, double x, double y, double z) {
    double rho_sqr = x * x + y * y + z * z;
    double w = rho_sqr - a * a;
    return 0.5 * (w + sqrt(w * w + 4.0 * a * a * z * z));
}
struct KsFK {
    double f, kx, ky, kz;
};
GD KsFK ks_f_k(double radius, double a, double x, double y, double z) {
    double r_sqr = ks_r_sqr(a, x, y, z);
    double r = sqrt(r_sqr);
    KsFK o;
    o.f = (r * r * r * radius) / (r * r * r * r + a * a * z * z);
    double den = r_sqr + a * a;
    o.kx = (r * x + a * y) / den;
    o.ky = (r * y - a * x) / den;
    o.kz = z / r;
    return o;
}
// covariant metric, symmetric: 10 unique entries g[0]=00 01 02 03 11 12 13 22 23 33
GD void ks_metric10(double radius, double a, double x, double y, double z, double g[10]) {
    KsFK k = ks_f_k(radius, a, x, y, z);
    double k0 = 1.0;
    g[0] = k0 * k0 * k.f - 1.0;
    g[1] = k0 * k.kx * k.f;
    g[2] = k0 * k.ky * k.f;
    g[3] = k0 * k.kz * k.f;
    g[4] = k.kx * k.kx * k.f + 1.0;
    g[5] = k.kx * k.ky * k.f;
    g[6] = k.kx * k.kz * k.f;
    g[7] = k.ky * k.ky * k.f + 1.0;
    g[8] = k.ky * k.kz * k.f;
    g[9] = k.kz * k.kz * k.f + 1.0;
}
GD void sym10_to_16(const double g[10], double m[4][4]) {
    m[0][0] = g[0]; m[0][1] = g[1]; m[0][2] = g[2]; m[0][3] = g[3];
    m[1][0] = g[1]; m[1][1] = g[4]; m[1][2] = g[5]; m[1][3] = g[6];
    m[2][0] = g[2]; m[2][1] = g[5]; m[2][2] = g[7]; m[2][3] = g[8];
    m[3][0] = g[3]; m[3][1] = g[6]; m[3][2] = g[8]; m[3][3] = g[9];
}
// contravariant metric: eta - (f k^i) k^j; NOT bitwise symmetric in the reference, all 16 kept
GD void ks_metric_contra(double radius, double a, double x, double y, double z, double m[4][4]) {
    KsFK k = ks_f_k(radius, a, x, y, z);
    double ku[4] = {-1.0, k.kx, k.ky, k.kz};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double eta = (i == j) ? (i == 0 ? -1.0 : 1.0) : 0.0;
            m[i][j] = eta - k.f * ku[i] * ku[j];
        }
}
GD void mat4_vec(const double m[4][4], const double v[4], double out[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = ((m[i][0] * v[0] + m[i][1] * v[1]) + m[i][2] * v[2]) + m[i][3] * v[3];
}
GD double vt_m_w(const double v[4], const double m[4][4], const double w[4]) {
    double row[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) row[j] = ((v[0] * m[0][j] + v[1] * m[1][j]) + v[2] * m[2][j]) + v[3] * m[3][j];
    return ((row[0] * w[0] + row[1] * w[1]) + row[2] * w[2]) + row[3] * w[3];
}

// ---- Boyer-Lindquist pieces (kerr_bl.rs:62-118, 253-272) ----------------------------------
GD double bl_delta(double r, double r_s, double a) { return r * r - r_s * r + a * a; }
struct BlMetric {
    double tt, rr, thth, phph, tph;
};
GD BlMetric bl_metric(double r_s, double a, double r, double sin_t, double cos_t) {
    double sig = r * r + a * a * (cos_t * cos_t);
    double sin2 = sin_t * sin_t;
    BlMetric g;
    g.tt = -(1.0 - r_s * r / sig);
    g.rr = sig / bl_delta(r, r_s, a);
    g.thth = sig;
    g.phph = (r * r + a * a + a * a * r_s * r * sin2 / sig) * sin2;
    g.tph = -a * r_s * r * sin2 / sig;
    return g;
}
// InnerProduct for KerrBL (kerr_bl.rs:339-358): the reference's mu-major double loop with the zero entries kept
GD double bl_inner(const BlMetric& g, const double v[4], const double w[4]) {
    double m[4][4] = {{g.tt, 0.0, 0.0, g.tph}, {0.0, g.rr, 0.0, 0.0}, {0.0, 0.0, g.thth, 0.0}, {g.tph, 0.0, 0.0, g.phph}};
    double result = 0.0;
#pragma unroll
    for (int mu = 0; mu < 4; ++mu)
#pragma unroll
        for (int nu = 0; nu < 4; ++nu) result += m[mu][nu] * v[mu] * w[nu];
    return result;
}

// ---- circular orbits (circular_orbit.rs:39-107) --------------------------------------------
GD double co_angular_velocity(double r_s, double a, double r) {
    double m = 0.5 * r_s;
    double sqrt_m = sqrt(m);
    return sqrt_m / (m_pow(r, 1.5) + a * sqrt_m);
}
GD bool co_killing(double r_s, double a, double r, double& u_t, double& u_phi) {
    double omega = co_angular_velocity(r_s, a, r);
    // metric_components_at(theta = pi/2): cos(pi/2)^2 and sin(pi/2)^2 as the reference evaluates them
    const double c = 6.123233995736766e-17;  // cos(FRAC_PI_2) in IEEE double (glibc and CUDA agree)
    double sig = r * r + a * a * (c * c);
    double sin2 = 1.0;
    double g_tt = -(1.0 - r_s * r / sig);
    double g_tphi = -a * r_s * r * sin2 / sig;
    double g_phiphi = (r * r + a * a + a * a * r_s * r * sin2 / sig) * sin2;
    double ut_pre = g_tt + 2.0 * omega * g_tphi + omega * omega * g_phiphi;
    if (ut_pre >= 0.0) return false;
    u_t = 1.0 / sqrt(-ut_pre);
    u_phi = omega * u_t;
    return true;
}

// ---- charts --------------------------------------------------------------------------------
enum { CH_CART = 0, CH_SPH = 1, CH_BL = 2 };
template <int G>
struct GeomTraits {
    static constexpr int chart = (G == GRRT_GEOM_EUCLIDEAN || G == GRRT_GEOM_KERR) ? CH_CART : (G == GRRT_GEOM_KERR_BL ? CH_BL : CH_SPH);
    static constexpr double sig0 = (G == GRRT_GEOM_KERR || G == GRRT_GEOM_KERR_BL) ? -1.0 : 1.0;
};

GD void cartesian_to_spherical(double x, double y, double z, double& r, double& theta, double& phi) {
    r = sqrt(x * x + y * y + z * z);
    if (r == 0.0) {
        theta = 0.0;
        phi = 0.0;
        return;
    }
    theta = m_acos(z / r);
    phi = m_atan2(y, x);
}
GD void cartesian_to_bl(double a, double x, double y, double z, double& r, double& theta, double& phi) {
    double rho_sqr = x * x + y * y + z * z;
    double w = rho_sqr - a * a;
    double r_sqr = 0.5 * (w + sqrt(w * w + 4.0 * a * a * z * z));
    r = sqrt(r_sqr);
    theta = (r == 0.0) ? 0.0 : m_acos(rclamp(z / r, -1.0, 1.0));
    phi = m_atan2(r * y - a * x, r * x + a * y);
}
// native position -> Cartesian embedding (point.rs:140-155, spherical_coordinates_helper.rs:28-39)
template <int G>
GD V3 position_cartesian(const DScene& S, const double* y) {
    if (GeomTraits<G>::chart == CH_CART) return v3(y[1], y[2], y[3]);
    double st, ct, sp, cp;
    sincos_pi2(y[2], &st, &ct);
    sincos_pi2(y[3], &sp, &cp);
    double r = y[1];
    if (GeomTraits<G>::chart == CH_SPH) return v3(r * st * cp, r * st * sp, r * ct);
    double a = S.a;
    return v3((r * cp - a * sp) * st, (r * sp + a * cp) * st, r * ct);
}
// Cartesian point -> native chart (point.rs:161-170)
template <int G>
GD void cartesian_to_native(const DScene& S, V3 p, double out[4]) {
    out[0] = 0.0;
    if (GeomTraits<G>::chart == CH_CART) {
        out[1] = p.x; out[2] = p.y; out[3] = p.z;
    } else if (GeomTraits<G>::chart == CH_SPH) {
        cartesian_to_spherical(p.x, p.y, p.z, out[1], out[2], out[3]);
    } else {
        cartesian_to_bl(S.a, p.x, p.y, p.z, out[1], out[2], out[3]);
    }
}

// ---- Geometry trait -------------------------------------------------------------------------
template <int G>
GD double inner_product(const DScene& S, const double* pos, const double* v, const double* w) {
    if (G == GRRT_GEOM_EUCLIDEAN) {
        return 1.0 * v[0] * w[0] + -v[1] * w[1] + -v[2] * w[2] + -v[3] * w[3];
    } else if (G == GRRT_GEOM_EUCLIDEAN_SPHERICAL) {
        double r = pos[1], s = m_sin(pos[2]);
        return 1.0 * v[0] * w[0] - v[1] * w[1] - r * r * v[2] * w[2] - r * r * s * s * v[3] * w[3];
    } else if (G == GRRT_GEOM_SCHWARZSCHILD) {
        double r = pos[1], s = m_sin(pos[2]);
        double a = 1.0 - S.radius / r;
        return a * v[0] * w[0] - v[1] * w[1] / a - r * r * v[2] * w[2] - r * r * s * s * v[3] * w[3];
    } else if (G == GRRT_GEOM_KERR) {
        double g10[10], m[4][4];
        ks_metric10(S.radius, S.a, pos[1], pos[2], pos[3], g10);
        sym10_to_16(g10, m);
        return vt_m_w(v, m, w);
    } else {
        double st, ct;
        m_sincos(pos[2], &st, &ct);
        BlMetric g = bl_metric(S.radius, S.a, pos[1], st, ct);
        return bl_inner(g, v, w);
    }
}
// get_radial_coordinate of a CARTESIAN point (disc.rs:112-120 calls it with the intersection point)
template <int G>
GD double radial_coordinate_cart(const DScene& S, V3 p) {
    if (G == GRRT_GEOM_KERR || G == GRRT_GEOM_KERR_BL) return sqrt(ks_r_sqr(S.a, p.x, p.y, p.z));
    return sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
}
template <int G>
GD void stationary_velocity(const DScene& S, const double* pos, double u[4]) {
    u[1] = u[2] = u[3] = 0.0;
    if (G == GRRT_GEOM_SCHWARZSCHILD) {
        double a = 1.0 - S.radius / pos[1];
        u[0] = 1.0 / sqrt(a);
    } else if (G == GRRT_GEOM_KERR) {
        KsFK k = ks_f_k(S.radius, S.a, pos[1], pos[2], pos[3]);
        u[0] = 1.0 / sqrt(1.0 - k.f);
    } else if (G == GRRT_GEOM_KERR_BL) {
        double c = m_cos(pos[2]);
        double sig = pos[1] * pos[1] + S.a * S.a * (c * c);
        u[0] = 1.0 / sqrt(1.0 - S.radius * pos[1] / sig);
    } else {
        u[0] = 1.0;
    }
}
// circular_orbit_killing_coefficients for a CARTESIAN sample point (volumetric_disc.rs:252-255)
template <int G>
GD bool orbit_coefficients_cart(const DScene& S, V3 p, double& u_t, double& u_phi) {
    if (G == GRRT_GEOM_EUCLIDEAN || G == GRRT_GEOM_EUCLIDEAN_SPHERICAL) {
        u_t = 1.0;
        u_phi = 0.0;
        return true;
    }
    double a = (G == GRRT_GEOM_SCHWARZSCHILD) ? 0.0 : S.a;
    return co_killing(S.radius, a, radial_coordinate_cart<G>(S, p), u_t, u_phi);
}
// get_circular_orbit_velocity_at for a point in the NATIVE chart
template <int G>
GD bool circular_orbit_velocity(const DScene& S, const double* pos, double u[4]) {
    if (G == GRRT_GEOM_EUCLIDEAN || G == GRRT_GEOM_EUCLIDEAN_SPHERICAL) {
        u[0] = 1.0; u[1] = u[2] = u[3] = 0.0;
        return true;
    }
    double u_t, u_phi;
    if (G == GRRT_GEOM_KERR) {
        double r = sqrt(ks_r_sqr(S.a, pos[1], pos[2], pos[3]));
        if (!co_killing(S.radius, S.a, r, u_t, u_phi)) return false;
        // u_t * (1,0,0,0) + u_phi * (0,-y,x,0)   (kerr.rs:473-485)
        u[0] = u_t * 1.0 + u_phi * 0.0;
        u[1] = u_t * 0.0 + u_phi * (-pos[2]);
        u[2] = u_t * 0.0 + u_phi * pos[1];
        u[3] = u_t * 0.0 + u_phi * 0.0;
        return true;
    }
    double a = (G == GRRT_GEOM_SCHWARZSCHILD) ? 0.0 : S.a;
    if (!co_killing(S.radius, a, pos[1], u_t, u_phi)) return false;
    u[0] = u_t; u[1] = 0.0; u[2] = 0.0; u[3] = u_phi;
    return true;
}

// ---- geodesic right-hand sides -----------------------------------------------------------------
struct RayConst {  // KerrBL per-ray constants of motion (kerr_bl.rs:120-126)
    double e, l_z, q;
};

// right-hand side with sin(theta), cos(theta) of the state supplied by the caller (spherical / Boyer-Lindquist
// charts read theta through them only; the Cartesian charts ignore them)
template <int G>
GD void geodesic_rhs_trig(const DScene& S, const RayConst& C, const double* y, double sin_theta, double cos_theta, double* d) {
    if (G == GRRT_GEOM_EUCLIDEAN) {  // euclidean.rs:46-53
        d[0] = y[4]; d[1] = y[5]; d[2] = y[6]; d[3] = y[7];
        d[4] = 0.0; d[5] = 0.0; d[6] = 0.0; d[7] = 0.0;
    } else if (G == GRRT_GEOM_EUCLIDEAN_SPHERICAL || G == GRRT_GEOM_SCHWARZSCHILD) {
        // schwarzschild.rs:54-80, euclidean_spherical.rs:46-70
        double r = y[1];
        double v_t = y[4], v_r = y[5], v_theta = y[6], v_phi = y[7];
        const double s = sin_theta, c = cos_theta;
#ifdef GRRT_FAST_MATH
        {   // same formulas with shared reciprocals (3 instead of 5-6 divisions)
            const double inv_r = frcp(r), inv_s = frcp(s);
            double a_t = 0.0, a_r;
            const double sph = v_theta * v_theta + v_phi * v_phi * s * s;
            if (G == GRRT_GEOM_SCHWARZSCHILD) {
                const double rr = S.radius * inv_r;
                const double a = 1.0 - rr;
                const double a_prime = rr * inv_r;
                const double apa = a_prime * frcp(a);
                a_t = -apa * v_t * v_r;
                a_r = -0.5 * a * a_prime * v_t * v_t + 0.5 * apa * v_r * v_r + a * r * sph;
            } else {
                a_r = r * sph;
            }
            const double two_over_r = 2.0 * inv_r;
            d[0] = v_t; d[1] = v_r; d[2] = v_theta; d[3] = v_phi;
            d[4] = a_t; d[5] = a_r;
            d[6] = -two_over_r * v_r * v_theta + s * c * v_phi * v_phi;
            d[7] = -two_over_r * v_phi * v_r - 2.0 * c * inv_s * v_theta * v_phi;
            return;
        }
#endif
        double a_t, a_r;
        if (G == GRRT_GEOM_SCHWARZSCHILD) {
            double a = 1.0 - S.radius / r;
            double a_prime = S.radius / (r * r);
            double apa = a_prime / a;
            a_t = -(apa)*v_t * v_r;
            a_r = -0.5 * a * a_prime * v_t * v_t + 0.5 * (apa)*v_r * v_r +
                  a * r * (v_theta * v_theta + v_phi * v_phi * s * s);
        } else {
            a_t = 0.0;
            a_r = r * (v_theta * v_theta + v_phi * v_phi * s * s);
        }
        double two_over_r = 2.0 / r;
        double a_theta = -(two_over_r)*v_r * v_theta + s * c * v_phi * v_phi;
        double a_phi = -(two_over_r)*v_phi * v_r - 2.0 * c / s * v_theta * v_phi;
        d[0] = v_t; d[1] = v_r; d[2] = v_theta; d[3] = v_phi;
        d[4] = a_t; d[5] = a_r; d[6] = a_theta; d[7] = a_phi;
    } else if (G == GRRT_GEOM_KERR_BL) {  // kerr_bl.rs:141-174
        double r = y[1], v_r = y[4], v_theta = y[5];
        double a = S.a, r_s = S.radius;
        const double sin_t = sin_theta, cos_t = cos_theta;
        double del = bl_delta(r, r_s, a);
        double rr_aa = r * r + a * a;
        double p_r = rr_aa * C.e - a * C.l_z;
        double sin2 = sin_t * sin_t;
#ifdef GRRT_FAST_MATH
        {   // same formulas with two shared reciprocals (1/Delta, 1/sin) instead of five divisions
            const double t1 = p_r * frcp(del);
            const double inv_s = frcp(sin_t), inv_s2 = inv_s * inv_s;
            const double ae = a * C.e;
            const double lae = C.l_z - ae;
            const double carter = lae * lae + C.q;
            d[0] = rr_aa * t1 + a * (C.l_z - ae * sin2);
            d[1] = v_r;
            d[2] = v_theta;
            d[3] = a * t1 + C.l_z * inv_s2 - ae;
            d[4] = 2.0 * r * C.e * p_r - (r - 0.5 * r_s) * carter;
            d[5] = cos_t * (C.l_z * C.l_z * inv_s2 * inv_s - ae * ae * sin_t);
            d[6] = 0.0; d[7] = 0.0;
            return;
        }
#endif
        double dt = rr_aa / del * p_r + a * (C.l_z - a * C.e * sin2);
        double dphi = a / del * p_r + C.l_z / sin2 - a * C.e;
        double lae = C.l_z - a * C.e;
        double carter = lae * lae + C.q;
        double dv_r = (4.0 * r * C.e * p_r - (2.0 * r - r_s) * carter) / 2.0;
        double dv_theta = (-2.0 * a * a * C.e * C.e * cos_t * sin_t + 2.0 * C.l_z * C.l_z * cos_t / ((sin_t * sin_t) * sin_t)) / 2.0;
        d[0] = dt; d[1] = v_r; d[2] = v_theta; d[3] = dphi;
        d[4] = dv_r; d[5] = dv_theta; d[6] = 0.0; d[7] = 0.0;
    } else {  // Kerr-Schild, kerr.rs:200-241
        double x = y[1], yy = y[2], z = y[3];
        double p[4] = {y[4], y[5], y[6], y[7]};
        double gc[4][4];
        ks_metric_contra(S.radius, S.a, x, yy, z, gc);
        double xdot[4];
        mat4_vec(gc, p, xdot);
        double acc[3];
        {
            acc[0] = acc[1] = acc[2] = 0.0;
            // the three axes run as a real loop: one copy of the ~700-instruction body instead of three keeps the
            // stage inside the instruction cache (the unrolled version stalled 1.56 cycles per issue on fetch)
#pragma unroll 1
            for (int idx = 0; idx < 3; ++idx) {
                // d_matrix_covariant (kerr.rs:162-186): central difference, h = 1e-10 * max(|x_i|, 1)
                const double xi = idx == 0 ? x : (idx == 1 ? yy : z);
                double h = 1e-10 * fmax(fabs(xi), 1.0);
                double dx = idx == 0 ? h : 0.0, dy = idx == 1 ? h : 0.0, dz = idx == 2 ? h : 0.0;
                double mp[10], mm[10], d10[10];
                ks_metric10(S.radius, S.a, x + dx, yy + dy, z + dz, mp);
                ks_metric10(S.radius, S.a, x - dx, yy - dy, z - dz, mm);
                double two_h = 2.0 * h;
#pragma unroll
                for (int k = 0; k < 10; ++k) d10[k] = (mp[k] - mm[k]) / two_h;
                double dm[4][4];
                sym10_to_16(d10, dm);
                // d_matrix_contravariant (kerr.rs:149-159): -(g^-1 * dg * g^-1), products accumulate k ascending
                double gd[4][4], t[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        gd[i][j] = ((gc[i][0] * dm[0][j] + gc[i][1] * dm[1][j]) + gc[i][2] * dm[2][j]) + gc[i][3] * dm[3][j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        t[i][j] = -(((gd[i][0] * gc[0][j] + gd[i][1] * gc[1][j]) + gd[i][2] * gc[2][j]) + gd[i][3] * gc[3][j]);
                const double ai = -0.5 * vt_m_w(p, t, p);
                if (idx == 0) acc[0] = ai;
                else if (idx == 1) acc[1] = ai;
                else acc[2] = ai;
            }
        }
        d[0] = xdot[0]; d[1] = xdot[1]; d[2] = xdot[2]; d[3] = xdot[3];
        d[4] = 0.0; d[5] = acc[0]; d[6] = acc[1]; d[7] = acc[2];
    }
}

template <int G>
GD void geodesic_rhs(const DScene& S, const RayConst& C, const double* y, double* d) {
    double s = 0.0, c = 1.0;
    if (GeomTraits<G>::chart != CH_CART) sincos_pi2(y[2], &s, &c);
    geodesic_rhs_trig<G>(S, C, y, s, c, d);
}

#ifdef GRRT_FAST_MATH
// sin and cos of a small angle, |x| <= 1/8: Taylor polynomials (truncation < 3e-18 for sin, 3e-20 for cos), no range
// reduction.  Used by the stage loop of the FMA build, where the polar angle of a stage argument differs from the
// angle of the state by the small increment sum B k.
__constant__ double c_trig_small[8] = {
    -1.66666666666666666667e-01, 8.33333333333333333333e-03, -1.98412698412698412698e-04, 2.75573192239858906526e-06,   // sin
    4.16666666666666666667e-02, -1.38888888888888888889e-03, 2.48015873015873015873e-05, -2.75573192239858906526e-07};  // cos
GD void sincos_small(double x, double* sn, double* cs) {
    const double z = x * x;
    double p = fma(z, c_trig_small[3], c_trig_small[2]);
    p = fma(z, p, c_trig_small[1]);
    p = fma(z, p, c_trig_small[0]);
    *sn = fma(x * z, p, x);
    double q = fma(z, c_trig_small[7], c_trig_small[6]);
    q = fma(z, q, c_trig_small[5]);
    q = fma(z, q, c_trig_small[4]);
    *cs = fma(z * z, q, fma(-0.5, z, 1.0));
}
#endif

// GeodesicSolver::momentum_from_state: contravariant momentum of a state
template <int G>
GD void momentum_from_state(const DScene& S, const RayConst& C, const double* y, double p[4]) {
    if (G == GRRT_GEOM_KERR) {  // kerr.rs:262-273
        double gc[4][4];
        ks_metric_contra(S.radius, S.a, y[1], y[2], y[3], gc);
        double pc[4] = {y[4], y[5], y[6], y[7]};
        mat4_vec(gc, pc, p);
    } else if (G == GRRT_GEOM_KERR_BL) {  // kerr_bl.rs:225-249
        double r = y[1], a = S.a;
        double st, ct;
        m_sincos(y[2], &st, &ct);
        double del = bl_delta(r, S.radius, a);
        double sig = r * r + a * a * (ct * ct);
        double sin2 = st * st;
        double rr_aa = r * r + a * a;
        double p_r_term = rr_aa * C.e - a * C.l_z;
        double dt = rr_aa / del * p_r_term + a * (C.l_z - a * C.e * sin2);
        double dphi = a / del * p_r_term + C.l_z / sin2 - a * C.e;
        p[0] = dt / sig; p[1] = y[4] / sig; p[2] = y[5] / sig; p[3] = dphi / sig;
    } else {
        p[0] = y[4]; p[1] = y[5]; p[2] = y[6]; p[3] = y[7];
    }
}

// Integrator::should_stop position tests (integrator.rs:203-268); `cart` is the Cartesian embedding of y
// The tests are evaluated branch-free and the first one that holds, in the reference's order, decides.
template <int G>
GD int stop_code(bool nan_pos, bool horizon, bool closed, bool escaped, bool nan_mom) {
    int st = nan_mom ? GRRT_STOP_NAN : GRRT_STOP_NONE;
    st = escaped ? GRRT_STOP_CELESTIAL : st;
    st = closed ? GRRT_STOP_CLOSED_ORBIT : st;
    st = horizon ? GRRT_STOP_HORIZON : st;
    st = nan_pos ? GRRT_STOP_NAN : st;
    return st;
}
template <int G>
GD void horizon_tests(const DScene& S, const double* y, unsigned long long step_index, bool& horizon, bool& closed) {
    horizon = false;
    closed = false;
    const bool last = step_index == S.max_steps - 1;
    if (G == GRRT_GEOM_SCHWARZSCHILD) {
        horizon = y[1] <= S.radius + S.heps;
        closed = last & (y[1] < 5.0 * S.radius);
    } else if (G == GRRT_GEOM_KERR_BL) {
        horizon = (S.horizon_enabled != 0) & (y[1] <= S.r_plus_eps);
        closed = last & (y[1] < 5.0 * S.radius);
    } else if (G == GRRT_GEOM_KERR) {
        const double r = sqrt(ks_r_sqr(S.a, y[1], y[2], y[3]));
        horizon = (S.horizon_enabled != 0) & (r <= S.r_plus_eps);
        closed = last & (r < 5.0 * S.radius);
    }
}
template <int G>
GD int should_stop(const DScene& S, const double* y, V3 cart, unsigned long long step_index) {
    bool horizon, closed;
    horizon_tests<G>(S, y, step_index, horizon, closed);
    const bool escaped = cart.x * cart.x + cart.y * cart.y + cart.z * cart.z > S.max_radius_sq;
    return stop_code<G>(!finite4(y), horizon, closed, escaped, !finite4(y + 4));
}

// ---- lazy geometry for the spherical / Boyer-Lindquist charts ------------------------------------------------------
// Converting a step to its Cartesian embedding costs two sincos, and almost no step can touch an object or the
// celestial sphere.  segment_needs_cartesian() decides from (r, theta) of the two end points alone, conservatively,
// whether the reference's tests on the chord could possibly succeed; if not, the conversion, the object loop and the
// escape test are skipped -- with identical results, because every skipped test is provably false:
//  * escape (integrator.rs:245-250): |x|^2 <= (r^2 + a^2)(1 + 4 eps) < max_radius^2;
//  * Disc (disc.rs:41-88): t = -z0/(z1-z0) lies in [0,1] only if z0, z1 have opposite signs (or |z1| < 2^-53 |z0|);
//    z = r cos(theta) has the sign of pi/2 - theta for theta in (0, pi), and the 1e-9 margin keeps |z1|/|z0| >> 2^-53;
//  * Sphere (sphere.rs:61-76): needs |x - c|^2 - R^2 to change sign; ||x| - |c|| > R at both ends rules that out;
//  * VolumetricDisc (volumetric_disc.rs:442-494): every capture surface lies in the slab |z| <= 3 thickness;
//    |z| >= r (2/pi) |theta - pi/2| (concavity of cos), so both ends beyond the slab on the same side cannot hit.
// Anything unusual (theta outside (0, pi), non-finite values, tilted disc axis) takes the full path.
template <int G>
GD bool segment_needs_cartesian(const DScene& S, double r0, double th0, double r1, double th1) {
    if (GeomTraits<G>::chart == CH_CART) return true;
    const double lo = 1e-9, hi = kPi - 1e-9;
    if (!(th0 > lo && th0 < hi && th1 > lo && th1 < hi && r0 > 0.0 && r1 > 0.0)) return true;
    const double aa = (GeomTraits<G>::chart == CH_BL) ? S.a * S.a : 0.0;
    const double q0 = r0 * r0, q1 = r1 * r1;
    if (!(q1 + aa < S.max_radius_sq * (1.0 - 1e-12))) return true;  // may have escaped: evaluate exactly
    const double s0 = th0 - 0.5 * kPi, s1 = th1 - 0.5 * kPi;        // sign(z) = -sign(s)
    const int flags = S.pre_flags;   // the per-object conditions of the comment above, and-ed over the objects at upload
    if (flags & PRE_ALWAYS) return true;
    if (flags & (PRE_HAS_DISC | PRE_HAS_VOL)) {
        if ((s0 > 0.0) != (s1 > 0.0)) return true;
        if ((flags & PRE_HAS_DISC) && !(fabs(s0) > 1e-9 && fabs(s1) > 1e-9)) return true;
        if (flags & PRE_HAS_VOL) {
            const double c = 2.0 / kPi, H = S.pre_vol_h;
            if (!(r0 * c * fabs(s0) > H && r1 * c * fabs(s1) > H)) return true;
        }
    }
    for (int k = 0; k < S.pre_n_spheres; ++k) {
        // |x| lies in [r, sqrt(r^2 + a^2)]; compared as squares so that no square root is needed
        const double far2 = S.pre_sph_far2[k], near2 = S.pre_sph_near2[k];
        const bool out0 = q0 > far2 || q0 + aa < near2;
        const bool out1 = q1 > far2 || q1 + aa < near2;
        if (!(out0 && out1)) return true;
    }
    return false;
}
// should_stop when segment_needs_cartesian() returned false (escape already excluded)
template <int G>
GD int should_stop_inside(const DScene& S, const double* y, unsigned long long step_index) {
    bool horizon, closed;
    horizon_tests<G>(S, y, step_index, horizon, closed);
    return stop_code<G>(!finite4(y), horizon, closed, false, !finite4(y + 4));
}

// ---- RKF45 tableau (runge_kutta.rs:16-84) ----------------------------------------------------------
#define RK_B21 (2.0 / 9.0)
#define RK_B31 (1.0 / 12.0)
#define RK_B32 (1.0 / 4.0)
#define RK_B41 (69.0 / 128.0)
#define RK_B42 (-243.0 / 128.0)
#define RK_B43 (135.0 / 64.0)
#define RK_B51 (-17.0 / 12.0)
#define RK_B52 (27.0 / 4.0)
#define RK_B53 (-27.0 / 5.0)
#define RK_B54 (16.0 / 15.0)
#define RK_B61 (65.0 / 432.0)
#define RK_B62 (-5.0 / 16.0)
#define RK_B63 (13.0 / 16.0)
#define RK_B64 (4.0 / 27.0)
#define RK_B65 (5.0 / 144.0)
#define RK_CH1 (47.0 / 450.0)
#define RK_CH2 0.0
#define RK_CH3 (12.0 / 25.0)
#define RK_CH4 (32.0 / 225.0)
#define RK_CH5 (1.0 / 30.0)
#define RK_CH6 (6.0 / 25.0)
#define RK_CT1 (1.0 / 150.0)
#define RK_CT2 0.0
#define RK_CT3 (-3.0 / 100.0)
#define RK_CT4 (16.0 / 75.0)
#define RK_CT5 (1.0 / 20.0)
#define RK_CT6 (-6.0 / 25.0)
constexpr double kHMax = 1.0, kHMin = 1e-12, kHGrowthCap = 4.0;

// one rkf45_step (runge_kutta.rs:86-125): y_new (5th order) and the truncation-error norm
template <int G>
GD void rkf45_step(const DScene& S, const RayConst& C, const double* y, double h, double* y_new, double& err) {
    double k1[8], k2[8], k3[8], k4[8], k5[8], k6[8], tmp[8];
    geodesic_rhs<G>(S, C, y, k1);
#pragma unroll
    for (int i = 0; i < 8; ++i) k1[i] = h * k1[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) tmp[i] = y[i] + RK_B21 * k1[i];
    geodesic_rhs<G>(S, C, tmp, k2);
#pragma unroll
    for (int i = 0; i < 8; ++i) k2[i] = h * k2[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) tmp[i] = (y[i] + RK_B31 * k1[i]) + RK_B32 * k2[i];
    geodesic_rhs<G>(S, C, tmp, k3);
#pragma unroll
    for (int i = 0; i < 8; ++i) k3[i] = h * k3[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) tmp[i] = ((y[i] + RK_B41 * k1[i]) + RK_B42 * k2[i]) + RK_B43 * k3[i];
    geodesic_rhs<G>(S, C, tmp, k4);
#pragma unroll
    for (int i = 0; i < 8; ++i) k4[i] = h * k4[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) tmp[i] = (((y[i] + RK_B51 * k1[i]) + RK_B52 * k2[i]) + RK_B53 * k3[i]) + RK_B54 * k4[i];
    geodesic_rhs<G>(S, C, tmp, k5);
#pragma unroll
    for (int i = 0; i < 8; ++i) k5[i] = h * k5[i];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        tmp[i] = ((((y[i] + RK_B61 * k1[i]) + RK_B62 * k2[i]) + RK_B63 * k3[i]) + RK_B64 * k4[i]) + RK_B65 * k5[i];
    geodesic_rhs<G>(S, C, tmp, k6);
#pragma unroll
    for (int i = 0; i < 8; ++i) k6[i] = h * k6[i];
    double e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        y_new[i] = (((((y[i] + RK_CH1 * k1[i]) + RK_CH2 * k2[i]) + RK_CH3 * k3[i]) + RK_CH4 * k4[i]) + RK_CH5 * k5[i]) + RK_CH6 * k6[i];
        double t = ((((RK_CT1 * k1[i] + RK_CT2 * k2[i]) + RK_CT3 * k3[i]) + RK_CT4 * k4[i]) + RK_CT5 * k5[i]) + RK_CT6 * k6[i];
        e[i] = t * t;
    }
    // nalgebra dotc with 8 accumulators: ((a0+a4)+(a1+a5))+(a2+a6))+(a3+a7)
    double res = (e[0] + e[4]);
    res = res + (e[1] + e[5]);
    res = res + (e[2] + e[6]);
    res = res + (e[3] + e[7]);
    err = sqrt(res);
}

// ---- colour / textures ----------------------------------------------------------------------------
GD Cie cie(double x, double y, double z, double a) { return Cie{x, y, z, a}; }
GD Cie blend(Cie self, Cie other) {  // color.rs:49-69
    double ab = rclamp(self.a, 0.0, 1.0);
    double af = rclamp(other.a, 0.0, 1.0);
    double ao = af + ab * (1.0 - af);
    if (ao <= 0.0) return cie(0.0, 0.0, 0.0, 0.0);
    double x = (other.x * af + self.x * ab * (1.0 - af)) / ao;
    double y = (other.y * af + self.y * ab * (1.0 - af)) / ao;
    double z = (other.z * af + self.z * ab * (1.0 - af)) / ao;
    return cie(x, y, z, ao);
}
GD Cie apply_beaming(Cie c, double redshift, double beaming) {  // color.rs:72-80
    double f = m_pow(redshift, beaming);
    return cie(c.x * f, c.y * f, c.z * f, c.a);
}
GD Cie texel_xyz(const DScene& S, const DTexture& t, uint32_t x, uint32_t y) {  // color.rs:32-40, 310-332
    const uchar4 px = __ldg(reinterpret_cast<const uchar4*>(t.rgba) + ((size_t)y * t.width + x));
    double r = __ldg(S.srgb_lut + px.x), g = __ldg(S.srgb_lut + px.y), b = __ldg(S.srgb_lut + px.z);
    double X = (0.4124564 * r + 0.3575761 * g) + 0.1804375 * b;
    double Y = (0.2126729 * r + 0.7151522 * g) + 0.0721750 * b;
    double Z = (0.0193339 * r + 0.1191920 * g) + 0.9503041 * b;
    return cie(X, Y, Z, (double)px.w / 255.0);
}
// out of line: one copy for the celestial sphere and every object texture, and none inside the ray-march loop
__device__ __noinline__ Cie bitmap_bilinear(const DScene& S, const DTexture& t, double u, double v) {  // texture.rs:62-90
    double p_x = (double)t.width * u;
    double p_y = (double)t.height * v;
    uint32_t xf = min(sat_u32(floor(p_x)), t.width - 1), yf = min(sat_u32(floor(p_y)), t.height - 1);
    uint32_t xc = min(sat_u32(ceil(p_x)), t.width - 1), yc = min(sat_u32(ceil(p_y)), t.height - 1);
    Cie c00 = texel_xyz(S, t, xf, yf), c01 = texel_xyz(S, t, xf, yc), c11 = texel_xyz(S, t, xc, yc), c10 = texel_xyz(S, t, xc, yf);
    double dx = p_x - (double)xf, dy = p_y - (double)yf;
    double w00 = (1.0 - dx) * (1.0 - dy), w01 = (1.0 - dx) * dy, w10 = dx * (1.0 - dy), w11 = dx * dy;
    Cie r;
    r.x = ((w00 * c00.x + w10 * c10.x) + w01 * c01.x) + w11 * c11.x;
    r.y = ((w00 * c00.y + w10 * c10.y) + w01 * c01.y) + w11 * c11.y;
    r.z = ((w00 * c00.z + w10 * c10.z) + w01 * c01.z) + w11 * c11.z;
    r.a = ((w00 * c00.a + w10 * c10.a) + w01 * c01.a) + w11 * c11.a;
    return r;
}
// index of the last key <= x in an ascending table (binary_search_by(total_cmp): Ok(i) => i, Err(i) => i-1);
// caller guarantees keys[0] < x < keys[n-1]
GD uint32_t lut_floor_index(const double* keys, uint32_t n, double x) {
    uint32_t lo = 0, hi = n;  // first index with key > x lies in (lo, hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}
GD Cie sample_blackbody(const DScene& S, double temperature) {  // texture.rs:149-194
    uint32_t n = S.bb_n;
    if (n == 0) return cie(0.0, 0.0, 0.0, 1.0);
    const double* lt = S.bb_log_t;
    const double* c = S.bb_xyz;
    double log_t = m_log10(fmax(temperature, 10.0));
    uint32_t lastk = n - 1;
    if (!isfinite(log_t) || log_t <= __ldg(lt)) return cie(__ldg(c), __ldg(c + 1), __ldg(c + 2), 1.0);
    if (log_t >= __ldg(lt + lastk)) return cie(__ldg(c + 3 * lastk), __ldg(c + 3 * lastk + 1), __ldg(c + 3 * lastk + 2), 1.0);
    uint32_t idx = lut_floor_index(lt, n, log_t);
    double lt0 = __ldg(lt + idx), lt1 = __ldg(lt + idx + 1);
    double t = (log_t - lt0) / (lt1 - lt0);
    const double* c0 = c + 3 * idx;
    double x0 = __ldg(c0), y0 = __ldg(c0 + 1), z0 = __ldg(c0 + 2), x1 = __ldg(c0 + 3), y1 = __ldg(c0 + 4), z1 = __ldg(c0 + 5);
    return cie(x0 + t * (x1 - x0), y0 + t * (y1 - y0), z0 + t * (z1 - z0), 1.0);
}
GD Cie texture_color(const DScene& S, const DTexture& t, double u, double v, double temperature, double redshift) {
    if (t.kind == GRRT_TEX_BITMAP) return apply_beaming(bitmap_bilinear(S, t, u, v), redshift, t.beaming);
    if (t.kind == GRRT_TEX_CHECKER) {  // texture.rs:240-256
        unsigned long long ut = sat_u64(floor(u * t.cw)), vt = sat_u64(floor(v * t.ch));
        const double* c = ((ut + vt) % 2ull == 0ull) ? t.c1 : t.c2;
        return apply_beaming(cie(c[0], c[1], c[2], c[3]), redshift, t.beaming);
    }
    return apply_beaming(sample_blackbody(S, temperature * redshift), redshift, t.beaming);  // texture.rs:198-210
}
// temperature.rs:197-253; returns GRRT_RAYERR_*
GD int compute_temperature(const DTemperature& tc, double radius, double& out) {
    if (tc.kind == GRRT_TEMP_CONSTANT) {
        out = tc.constant;
        return 0;
    }
    if (!isfinite(radius)) return GRRT_RAYERR_NONFINITE_RADIUS;
    if (radius < tc.r_isco || tc.n == 0) return GRRT_RAYERR_BELOW_RISCO;
    uint32_t lastk = tc.n - 1;
    if (radius <= __ldg(tc.lut_r)) {
        out = __ldg(tc.lut_t);
        return 0;
    }
    if (radius >= __ldg(tc.lut_r + lastk)) {
        out = __ldg(tc.lut_t + lastk);
        return 0;
    }
    uint32_t idx = lut_floor_index(tc.lut_r, tc.n, radius);
    double r0 = __ldg(tc.lut_r + idx), r1 = __ldg(tc.lut_r + idx + 1);
    double t0 = __ldg(tc.lut_t + idx), t1 = __ldg(tc.lut_t + idx + 1);
    double t = (radius - r0) / (r1 - r0);
    out = t0 + t * (t1 - t0);
    return 0;
}

// ---- scene objects: hit tests on the straight chord between two steps -------------------------------
struct Hit {
    double u, v, t;
    V3 point, dir;
};

GD bool disc_intersects(const DObject& o, V3 ys, V3 ye, Hit& h) {  // disc.rs:41-88
    V3 d = ye - ys;
    // (center - y_start).dot(normal) with normal = (0,0,1), center = 0: ((0-x)*0 + (0-y)*0) + (0-z)*1
    double p1 = ((0.0 - ys.x) * 0.0 + (0.0 - ys.y) * 0.0) + (0.0 - ys.z) * 1.0;
    double p2 = (d.x * 0.0 + d.y * 0.0) + d.z * 1.0;
    double t = p1 / p2;
    if (!in01(t)) return false;
    V3 ip = ys + t * d;
    double rr = dot3(ip, ip);
    if (rr >= o.r_in * o.r_in && rr <= o.r_out * o.r_out) {
        double phi = m_atan2(ip.y - 0.0, ip.x - 0.0);
        double r = (sqrt(rr) - o.r_in) / (o.r_out - o.r_in);
        double sp, cp;
        m_sincos(phi, &sp, &cp);
        h.u = 0.5 + 0.5 * r * cp;
        h.v = 0.5 + 0.5 * r * sp;
        h.point = ip;
        h.t = t;
        h.dir = d;
        return true;
    }
    return false;
}
GD bool sphere_intersects(const DObject& o, V3 ys_c, V3 ye_c, Hit& h) {  // sphere.rs:37-127
    V3 neg = v3(-o.pos[0], -o.pos[1], -o.pos[2]);
    V3 ys = ys_c + neg, ye = ye_c + neg;
    double r_start = ys.x * ys.x + ys.y * ys.y + ys.z * ys.z;
    double r_end = ye.x * ye.x + ye.y * ye.y + ye.z * ye.z;
    double R2 = o.radius * o.radius;
    if (!((r_start >= R2 && r_end <= R2) || (r_start <= R2 && r_end >= R2))) return false;
    V3 d = ye - ys;
    double a = dot3(d, d);
    double b = 2.0 * dot3(ys, d);
    double c = dot3(ys, ys) - o.radius * o.radius;
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return false;
    double sq = sqrt(disc);
    double t1 = (-b + sq) / (2.0 * a), t2 = (-b - sq) / (2.0 * a);
    double t;
    if (in01(t1)) t = t1;
    else if (in01(t2)) t = t2;
    else return false;
    V3 ps = ys + t * d;
    double r, theta, phi;
    cartesian_to_spherical(ps.x, ps.y, ps.z, r, theta, phi);
    double u = (kPi + phi) / (2.0 * kPi);
    h.u = 1.0 - u;
    h.v = theta / kPi;
    h.point = v3(ps.x + o.pos[0], ps.y + o.pos[1], ps.z + o.pos[2]);
    h.t = t;
    h.dir = d;
    return true;
}

// --- VolumetricDisc (volumetric_disc.rs) ---
constexpr double kMinIntersectionT = 1e-9;
GD V3 obj_axis(const DObject& o) { return v3(o.axis[0], o.axis[1], o.axis[2]); }

// intersects_clipped_cylinder (:348-405): appends 0..2 hits (already min/max ordered)
GD void vol_clipped_cylinder(V3 from, V3 to, double radius, V3 axis, double half_height, double* hits, int& n) {
    V3 seg = to - from;
    double seg_len = norm3(seg);
    if (seg_len < 1e-12) return;
    V3 d = seg / seg_len;
    V3 v = cross3(from, axis), w = cross3(d, axis);
    double a = dot3(w, w);
    double b = 2.0 * dot3(v, w);
    double c = dot3(v, v) - radius * radius;
    if (a < 1e-10) return;  // NoIntersection or Parallel: neither contributes a hit
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return;
    double sq = sqrt(disc);
    double dist[2] = {(-b - sq) / (2.0 * a), (-b + sq) / (2.0 * a)};
    double hh[2];
    int m = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        double t = dist[k] / seg_len;
        if (in01(t)) {
            V3 p = from + t * (to - from);
            if (fabs(dot3(p, axis)) <= half_height) hh[m++] = t;
        }
    }
    if (m == 1) hits[n++] = hh[0];
    if (m == 2) {
        hits[n++] = fmin(hh[0], hh[1]);
        hits[n++] = fmax(hh[0], hh[1]);
    }
}
GD bool vol_cap(V3 from, V3 to, double radius, V3 axis, double pos, double& t_out) {  // :407-440
    V3 seg = to - from;
    double seg_len = norm3(seg);
    if (seg_len < 1e-12) return false;
    V3 n = seg / seg_len;
    double denom = dot3(n, axis);
    if (fabs(denom) < 1e-10) return false;
    double t = (pos - dot3(from, axis)) / dot3(seg, axis);
    if (!in01(t)) return false;
    V3 p = from + t * seg;
    V3 cr = cross3(p, axis);
    if (dot3(cr, cr) > radius * radius) return false;
    t_out = t;
    return true;
}
// intersects_cylinder (:442-494) followed by the "first hit with t > 1e-9" selection used by both
// intersects (:530-549) and precompute_exit_distance (:183-195).  Returns false when there is none.
GD bool vol_first_hit(const DObject& o, V3 from, V3 to, double& t_sel) {
    double hits[6];
    int n = 0;
    V3 axis = obj_axis(o);
    V3 dir = to - from;
    double capture_height = o.thickness * 3.0;
    vol_clipped_cylinder(from, to, o.r_out, axis, capture_height, hits, n);
    vol_clipped_cylinder(from, to, o.r_in, axis, capture_height, hits, n);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        double pos = k == 0 ? capture_height : -capture_height;
        double t;
        if (vol_cap(from, to, o.r_out, axis, pos, t)) {
            V3 p = from + t * dir;
            V3 cr = cross3(p, axis);
            if (dot3(cr, cr) >= o.r_in * o.r_in) hits[n++] = t;
        }
    }
    if (n == 0) return false;
    // two smallest (all hits are finite, non-negative: total_cmp order == numeric order)
    double m1 = CUDART_INF, m2 = CUDART_INF;
    for (int k = 0; k < n; ++k) {
        double t = hits[k];
        if (t < m1) { m2 = m1; m1 = t; }
        else if (t < m2) m2 = t;
    }
    if (m1 > kMinIntersectionT) { t_sel = m1; return true; }
    if (n >= 2 && m2 > kMinIntersectionT) { t_sel = m2; return true; }
    return false;
}
GD void vol_get_uv(const DObject& o, V3 p, double& u, double& v) {  // :140-152
    double x = dot3(p, v3(o.e1[0], o.e1[1], o.e1[2])), y = dot3(p, v3(o.e2[0], o.e2[1], o.e2[2]));
    double rr = sqrt(x * x + y * y);
    double phi = m_atan2(y, x);
    double r = (rr - o.r_in) / (o.r_out - o.r_in);
    double sp, cp;
    m_sincos(phi, &sp, &cp);
    u = 0.5 + 0.5 * r * cp;
    v = 0.5 + 0.5 * r * sp;
}
GD bool vol_intersects(const DObject& o, V3 ys, V3 ye, Hit& h) {  // :506-578
    double t;
    if (!vol_first_hit(o, ys, ye, t)) return false;
    if (!in01(t)) return false;
    V3 d = ye - ys;
    V3 ip = ys + t * d;
    vol_get_uv(o, ip, h.u, h.v);
    h.point = ip;
    h.t = t;
    h.dir = d;
    return true;
}

// Perlin noise of crate noise 0.9.0 (core/perlin.rs perlin_3d); perm = 256-byte table in shared memory
GD unsigned perlin_hash(const uint8_t* perm, long long x, long long y, long long z) {
    unsigned a = (unsigned)(x & 0xff), b = (unsigned)(y & 0xff), c = (unsigned)(z & 0xff);
    return perm[perm[perm[a] ^ b] ^ c];
}
// gradient dot product of noise 0.9.0 (core/perlin.rs: `match hash & 15`): one of (+-x +-y), (+-x +-z), (+-y +-z),
// cases 12..15 repeating (x+y), (-x+y), (-y+z), (-y-z).  Branch-free -- a switch diverges 12 ways inside a warp --
// with the same IEEE result: operands are selected, their sign bits flipped, then added once.
GD double perlin_grad(unsigned hsh, double x, double y, double z) {
    const unsigned h = hsh & 15u;
    const double u = ((0xCF00u >> h) & 1u) ? y : x;   // first operand is y for 8..11, 14, 15
    const double v = ((0xCFF0u >> h) & 1u) ? z : y;   // second operand is z for 4..11, 14, 15
    const unsigned long long su = (unsigned long long)((0xEAAAu >> h) & 1u) << 63;  // first operand negated
    const unsigned long long sv = (unsigned long long)((0x8CCCu >> h) & 1u) << 63;  // second operand negated
    return __longlong_as_double(__double_as_longlong(u) ^ (long long)su) + __longlong_as_double(__double_as_longlong(v) ^ (long long)sv);
}
GD double quintic(double x) { return x * x * x * (x * (x * 6.0 - 15.0) + 10.0); }
GD double perlin_3d(const uint8_t* perm, double px, double py, double pz) {
    double fx = floor(px), fy = floor(py), fz = floor(pz);
    long long cx = (long long)fx, cy = (long long)fy, cz = (long long)fz;
    double dx = px - fx, dy = py - fy, dz = pz - fz;
    double g000 = perlin_grad(perlin_hash(perm, cx, cy, cz), dx, dy, dz);
    double g100 = perlin_grad(perlin_hash(perm, cx + 1, cy, cz), dx - 1.0, dy, dz);
    double g010 = perlin_grad(perlin_hash(perm, cx, cy + 1, cz), dx, dy - 1.0, dz);
    double g110 = perlin_grad(perlin_hash(perm, cx + 1, cy + 1, cz), dx - 1.0, dy - 1.0, dz);
    double g001 = perlin_grad(perlin_hash(perm, cx, cy, cz + 1), dx, dy, dz - 1.0);
    double g101 = perlin_grad(perlin_hash(perm, cx + 1, cy, cz + 1), dx - 1.0, dy, dz - 1.0);
    double g011 = perlin_grad(perlin_hash(perm, cx, cy + 1, cz + 1), dx, dy - 1.0, dz - 1.0);
    double g111 = perlin_grad(perlin_hash(perm, cx + 1, cy + 1, cz + 1), dx - 1.0, dy - 1.0, dz - 1.0);
    double a = quintic(dx), b = quintic(dy), c = quintic(dz);
    double k0 = g000;
    double k1 = g100 - g000;
    double k2 = g010 - g000;
    double k3 = g001 - g000;
    double k4 = g000 + g110 - g100 - g010;
    double k5 = g000 + g101 - g100 - g001;
    double k6 = g000 + g011 - g010 - g001;
    double k7 = g100 + g010 + g001 + g111 - g000 - g110 - g101 - g011;
    double unscaled = k0 + k1 * a + k2 * b + k3 * c + k4 * a * b + k5 * a * c + k6 * b * c + k7 * a * b * c;
    return rclamp(unscaled * 1.1547005383792515, -1.0, 1.0);
}
GD double vol_density(const DObject& o, const uint8_t* perm, V3 p) {  // :97-138
    V3 axis = obj_axis(o);
    double h = fabs(dot3(p, axis));
    double r = norm3(cross3(p, axis));
    if (r <= o.r_in || r >= o.r_out) return 0.0;
    double ht = h / o.thickness;
    double vertical_falloff = m_exp(-(ht * ht));
    if (vertical_falloff < 0.001) return 0.0;
    double radial_base = m_pow(o.r_in / r, 1.5);
    double boundary_falloff = 1.0;
    double d1 = o.r_out - r, d2 = r - o.r_in;
    boundary_falloff *= m_exp(-1.0 / fmax(d1 * d1, 0.0001));
    boundary_falloff *= m_exp(-1.0 / fmax(d2 * d2, 0.0001));
    double x_local = dot3(p, v3(o.e1[0], o.e1[1], o.e1[2])), y_local = dot3(p, v3(o.e2[0], o.e2[1], o.e2[2]));
    double phi = m_atan2(y_local, x_local);
    double sphi, cphi;
    m_sincos(phi, &sphi, &cphi);
    double npx = r * o.noise_scale[0], npy = cphi * o.noise_scale[1], npz = sphi * o.noise_scale[1];
    // fbm(noise_p, 0.5) (:330-342): g = exp2(-0.5)
    const double g = 0.70710678118654757;
    double frequency = 4.0, amplitude = 1.0, n = 0.0;
    for (unsigned long long k = 0; k < o.num_octaves; ++k) {
        n += amplitude * perlin_3d(perm, npx * frequency, npy * frequency, npz * frequency);
        frequency *= 2.0;
        amplitude *= g;
    }
    n += perlin_3d(perm, r * 0.5, h * o.noise_scale[2], cphi) * 0.5;
    double nn = fmax(n + o.noise_offset, 0.0) * o.density_multiplier;
    return nn * radial_base * vertical_falloff * boundary_falloff;
}

// ---- adaptive sampling hash (raytracer.rs:132-159) -----------------------------------------------
GD unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
GD double hash_pixel_samples(long long row, long long col, unsigned long long k) {
    unsigned long long z = mix64((unsigned long long)row + mix64((unsigned long long)col + mix64(k)));
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace GRRT_NS
