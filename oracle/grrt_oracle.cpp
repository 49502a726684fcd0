// grrt_oracle.cpp -- CPU ORACLE for gr_raytracer's per-pixel geodesic path.
//
// TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement of the reference's algorithm
// (mdreem/gr_raytracer, Rust; paths below are relative to /root/reference).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may build,
// load or call it.  The product (libgrrt.so) never links it and has no CPU fallback.
//
// Build: oracle/Makefile  (g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp).  No FMA
// contraction and glibc libm, i.e. the same sin/cos/pow/exp/atan2/acos/log10 a Rust
// linux-gnu build calls, so results differ from the reference only through nalgebra's
// summation order, which is restated below from nalgebra 0.35's blas.rs (see A.9 of
// SURVEY.md): Matrix*Vector and Matrix*Matrix accumulate k ascending; Vector3 dot is
// (a+b)+c; the 8-component norm is sqrt((((a0+a4)+(a1+a5))+(a2+a6))+(a3+a7)).
//
// Parity pinning: the reference cannot be compiled here (no cargo/rustc), so this oracle
// is pinned against every known-answer value the reference's own unit tests hold for the
// path (tests/test_oracle_kat.py).  The Perlin noise of VolumetricDisc
// comes from the un-vendored crate `noise 0.9.0` (+ rand 0.8 / rand_xorshift 0.3); its
// published algorithm is restated in `perlin_*` below.  No reference UNIT test pins a noise
// value; the restatement is pinned instead against the image the reference binary itself
// rendered with a frozen recipe (kerr_black_hole_with_stars.png, images/create-main-image.sh:
// 40-79): tests/test_reference_image.py compares the oracle's render of that recipe with
// crops of the image (a wrong permutation table fails it by a wide margin).
//
// Each function cites the reference file:line it follows.

#include "../include/grrt.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double PI = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------
// small linear algebra with nalgebra's evaluation order
// ---------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
struct V4 {
    double v[4];
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
struct M4 {
    double m[4][4];
};
struct Y8 {
    double v[8];
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};

inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
inline double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double norm_sq(V3 a) { return dot(a, a); }
inline double norm(V3 a) { return std::sqrt(norm_sq(a)); }
inline V3 normalize(V3 a) { return a / norm(a); }

inline V4 v4(double a, double b, double c, double d) { return V4{{a, b, c, d}}; }
inline V4 operator+(V4 a, V4 b) { return v4(a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]); }
inline V4 operator*(double s, V4 a) { return v4(s * a[0], s * a[1], s * a[2], s * a[3]); }
inline V4 operator-(V4 a) { return v4(-a[0], -a[1], -a[2], -a[3]); }
inline V4 operator/(V4 a, double s) { return v4(a[0] / s, a[1] / s, a[2] / s, a[3] / s); }

inline V4 mul(const M4& a, V4 x) {  // gemv: columns accumulated k ascending
    V4 r;
    for (int i = 0; i < 4; ++i) {
        double acc = a.m[i][0] * x[0];
        acc = acc + a.m[i][1] * x[1];
        acc = acc + a.m[i][2] * x[2];
        acc = acc + a.m[i][3] * x[3];
        r[i] = acc;
    }
    return r;
}
inline M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = a.m[i][0] * b.m[0][j];
            acc = acc + a.m[i][1] * b.m[1][j];
            acc = acc + a.m[i][2] * b.m[2][j];
            acc = acc + a.m[i][3] * b.m[3][j];
            r.m[i][j] = acc;
        }
    return r;
}
inline double vt_m_w(V4 v, const M4& g, V4 w) {  // (v^T * g) * w
    double row[4];
    for (int j = 0; j < 4; ++j) {
        double acc = v[0] * g.m[0][j];
        acc = acc + v[1] * g.m[1][j];
        acc = acc + v[2] * g.m[2][j];
        acc = acc + v[3] * g.m[3][j];
        row[j] = acc;
    }
    double acc = row[0] * w[0];
    acc = acc + row[1] * w[1];
    acc = acc + row[2] * w[2];
    acc = acc + row[3] * w[3];
    return acc;
}
inline M4 zeros4() {
    M4 r;
    std::memset(&r, 0, sizeof r);
    return r;
}

// Rust semantics helpers
inline double rclamp(double x, double lo, double hi) {  // f64::clamp
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
inline double rmax(double a, double b) { return std::fmax(a, b); }  // f64::max (NaN -> other)
inline double rmin(double a, double b) { return std::fmin(a, b); }
inline double rem_euclid(double a, double b) {
    double r = std::fmod(a, b);
    return r < 0.0 ? r + std::fabs(b) : r;
}
inline uint32_t sat_u32(double x) {  // `as u32`
    if (!(x > 0.0)) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}
inline uint64_t sat_u64(double x) {  // `as usize`
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)x;
}
inline bool in01(double t) { return t >= 0.0 && t <= 1.0; }  // (0.0..=1.0).contains
inline double powi2(double x) { return x * x; }
inline double powi3(double x) { return (x * x) * x; }
inline double powi4(double x) {
    double s = x * x;
    return s * s;
}
inline double powi5(double x) {
    double s = x * x;
    return (s * s) * x;
}

// ---------------------------------------------------------------------------------------
// charts  (geometry/point.rs, geometry/spherical_coordinates_helper.rs)
// ---------------------------------------------------------------------------------------
enum Chart { CART = 0, SPH = 1, BL = 2 };

struct Geom {
    int kind;
    double radius, a, heps;
    int kerr_mode;
    Chart chart() const {
        switch (kind) {
            case GRRT_GEOM_EUCLIDEAN:
            case GRRT_GEOM_KERR: return CART;
            case GRRT_GEOM_KERR_BL: return BL;
            default: return SPH;
        }
    }
    double sig0() const {  // signature()[0]
        return (kind == GRRT_GEOM_KERR || kind == GRRT_GEOM_KERR_BL) ? -1.0 : 1.0;
    }
    double sig3() const { return -sig0(); }
};

// spherical_coordinates_helper.rs:5-26
V4 cartesian_to_spherical(V4 c) {
    double t = c[0], x = c[1], y = c[2], z = c[3];
    double r = std::sqrt(x * x + y * y + z * z);
    if (r == 0.0) return v4(t, 0.0, 0.0, 0.0);
    double theta = std::acos(z / r);
    double phi = std::atan2(y, x);
    return v4(t, r, theta, phi);
}
// spherical_coordinates_helper.rs:28-39
V4 spherical_to_cartesian(V4 s) {
    double t = s[0], r = s[1], theta = s[2], phi = s[3];
    double x = r * std::sin(theta) * std::cos(phi);
    double y = r * std::sin(theta) * std::sin(phi);
    double z = r * std::cos(theta);
    return v4(t, x, y, z);
}
// spherical_coordinates_helper.rs:44-61
V4 cartesian_to_boyer_lindquist(double a, V4 c) {
    double t = c[0], x = c[1], y = c[2], z = c[3];
    double rho_sqr = x * x + y * y + z * z;
    double r_sqr = 0.5 * (rho_sqr - a * a + std::sqrt(powi2(rho_sqr - a * a) + 4.0 * a * a * z * z));
    double r = std::sqrt(r_sqr);
    double theta = (r == 0.0) ? 0.0 : std::acos(rclamp(z / r, -1.0, 1.0));
    double phi = std::atan2(r * y - a * x, r * x + a * y);
    return v4(t, r, theta, phi);
}
// point.rs:140-155 to_cartesian
V4 bl_to_cartesian(double a, V4 p) {
    double t = p[0], r = p[1], theta = p[2], phi = p[3];
    double x = (r * std::cos(phi) - a * std::sin(phi)) * std::sin(theta);
    double y = (r * std::sin(phi) + a * std::cos(phi)) * std::sin(theta);
    double z = r * std::cos(theta);
    return v4(t, x, y, z);
}
V4 to_cartesian(const Geom& g, Chart c, V4 p) {
    switch (c) {
        case CART: return p;
        case SPH: return spherical_to_cartesian(p);
        default: return bl_to_cartesian(g.a, p);
    }
}
// point.rs:161-170 to_coordinate_system (source chart is always Cartesian on this path)
V4 cartesian_to_chart(const Geom& g, Chart target, V4 p) {
    switch (target) {
        case CART: return p;
        case SPH: return cartesian_to_spherical(p);
        default: return cartesian_to_boyer_lindquist(g.a, p);
    }
}
double wrap_theta(double th) { return rem_euclid(th, PI); }                     // point.rs:79-81
double wrap_phi(double ph) { return rem_euclid(ph + PI, 2.0 * PI) - PI; }       // point.rs:84-86
// point.rs:173-188 get_as_spherical -> (r, theta, phi)
V3 get_as_spherical(Chart c, V4 p) {
    if (c == CART) {
        V4 s = cartesian_to_spherical(p);
        return v3(s[1], s[2], s[3]);
    }
    return v3(p[1], wrap_theta(p[2]), wrap_phi(p[3]));
}
V3 spatial_cartesian(const Geom& g, Chart c, V4 p) {  // point.rs:125-138
    V4 q = to_cartesian(g, c, p);
    return v3(q[1], q[2], q[3]);
}

// ---------------------------------------------------------------------------------------
// Kerr-Schild metric (geometry/kerr.rs:31-110)
// ---------------------------------------------------------------------------------------
double ks_r_sqr(double a, double x, double y, double z) {  // kerr.rs:31-34
    double rho_sqr = x * x + y * y + z * z;
    return 0.5 * (rho_sqr - a * a +
                  std::sqrt((rho_sqr - a * a) * (rho_sqr - a * a) + 4.0 * a * a * z * z));
}
V4 ks_k_covector(double a, double x, double y, double z) {  // kerr.rs:36-46
    double r_sqr = ks_r_sqr(a, x, y, z);
    double r = std::sqrt(r_sqr);
    double k_x = (r * x + a * y) / (r_sqr + a * a);
    double k_y = (r * y - a * x) / (r_sqr + a * a);
    double k_z = z / r;
    return v4(1.0, k_x, k_y, k_z);
}
double ks_f(double radius, double a, double x, double y, double z) {
    double r_sqr = ks_r_sqr(a, x, y, z);
    double r = std::sqrt(r_sqr);
    return (r * r * r * radius) / (r * r * r * r + a * a * z * z);
}
M4 ks_metric(double radius, double a, double x, double y, double z) {  // kerr.rs:49-84
    double f = ks_f(radius, a, x, y, z);
    V4 k = ks_k_covector(a, x, y, z);
    double k_0 = k[0], k_x = k[1], k_y = k[2], k_z = k[3];
    M4 g = zeros4();
    g.m[0][0] = k_0 * k_0 * f - 1.0;
    g.m[0][1] = k_0 * k_x * f;
    g.m[0][2] = k_0 * k_y * f;
    g.m[0][3] = k_0 * k_z * f;
    g.m[1][0] = g.m[0][1];
    g.m[1][1] = k_x * k_x * f + 1.0;
    g.m[1][2] = k_x * k_y * f;
    g.m[1][3] = k_x * k_z * f;
    g.m[2][0] = g.m[0][2];
    g.m[2][1] = g.m[1][2];
    g.m[2][2] = k_y * k_y * f + 1.0;
    g.m[2][3] = k_y * k_z * f;
    g.m[3][0] = g.m[0][3];
    g.m[3][1] = g.m[1][3];
    g.m[3][2] = g.m[2][3];
    g.m[3][3] = k_z * k_z * f + 1.0;
    return g;
}
M4 ks_metric_contravariant(double radius, double a, double x, double y, double z) {  // kerr.rs:88-110
    double f = ks_f(radius, a, x, y, z);
    V4 kc = ks_k_covector(a, x, y, z);
    double ku[4] = {-kc[0], kc[1], kc[2], kc[3]};
    M4 g = zeros4();
    g.m[0][0] = -1.0;
    g.m[1][1] = 1.0;
    g.m[2][2] = 1.0;
    g.m[3][3] = 1.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) g.m[i][j] -= f * ku[i] * ku[j];
    return g;
}
// kerr.rs:162-186: central finite difference of the covariant metric
M4 ks_d_metric_cov_fd(double radius, double a, int index, double x, double y, double z) {
    if (index == 0) return zeros4();
    double base = 1e-10;
    double h = base * (index == 1 ? rmax(std::fabs(x), 1.0)
                                  : index == 2 ? rmax(std::fabs(y), 1.0) : rmax(std::fabs(z), 1.0));
    double dx = index == 1 ? h : 0.0, dy = index == 2 ? h : 0.0, dz = index == 3 ? h : 0.0;
    M4 mp = ks_metric(radius, a, x + dx, y + dy, z + dz);
    M4 mm = ks_metric(radius, a, x - dx, y - dy, z - dz);
    M4 d;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) d.m[i][j] = (mp.m[i][j] - mm.m[i][j]) / (2.0 * h);
    return d;
}

// ---------------------------------------------------------------------------------------
// Boyer-Lindquist pieces (geometry/kerr_bl.rs:62-118, 253-272)
// ---------------------------------------------------------------------------------------
const double SIN2_FLOOR = 1e-28;  // kerr_bl.rs:28
double bl_sigma(double r, double a, double theta) { return r * r + a * a * powi2(std::cos(theta)); }
double bl_delta(double r, double r_s, double a) { return r * r - r_s * r + a * a; }
M4 bl_metric(double r_s, double a, double r, double theta) {  // kerr_bl.rs:253-272
    double sig = bl_sigma(r, a, theta);
    double sin_t = std::sin(theta);
    double sin2 = sin_t * sin_t;
    M4 g = zeros4();
    g.m[0][0] = -(1.0 - r_s * r / sig);
    g.m[1][1] = sig / bl_delta(r, r_s, a);
    g.m[2][2] = sig;
    g.m[3][3] = (r * r + a * a + a * a * r_s * r * sin2 / sig) * sin2;
    g.m[0][3] = -a * r_s * r * sin2 / sig;
    g.m[3][0] = g.m[0][3];
    return g;
}
double bl_potential_r(double r, double r_s, double a, double e, double l_z, double q) {  // :78-82
    double del = bl_delta(r, r_s, a);
    double p_r = (r * r + a * a) * e - a * l_z;
    return p_r * p_r - del * (powi2(l_z - a * e) + q);
}
double bl_potential_r_derivative(double r, double r_s, double a, double e, double l_z, double q) {  // :85-89
    double p_r = (r * r + a * a) * e - a * l_z;
    double carter_term = powi2(l_z - a * e) + q;
    return 4.0 * r * e * p_r - (2.0 * r - r_s) * carter_term;
}
double bl_potential_theta(double theta, double a, double e, double l_z, double q) {  // :101-105
    double cos_t = std::cos(theta), sin_t = std::sin(theta);
    return q + a * a * e * e * cos_t * cos_t - l_z * l_z * cos_t * cos_t / (sin_t * sin_t);
}
double bl_potential_theta_derivative(double theta, double a, double e, double l_z) {  // :114-118
    double cos_t = std::cos(theta), sin_t = std::sin(theta);
    return -2.0 * a * a * e * e * cos_t * sin_t + 2.0 * l_z * l_z * cos_t / powi3(sin_t);
}

// ---------------------------------------------------------------------------------------
// circular orbits (geometry/circular_orbit.rs:39-136)
// ---------------------------------------------------------------------------------------
void co_metric_components_at(double r_s, double a, double r, double theta, double& g_tt,
                             double& g_tphi, double& g_phiphi) {  // :39-46
    double sig = r * r + a * a * powi2(std::cos(theta));
    double sin2 = powi2(std::sin(theta));
    g_tt = -(1.0 - r_s * r / sig);
    g_tphi = -a * r_s * r * sin2 / sig;
    g_phiphi = (r * r + a * a + a * a * r_s * r * sin2 / sig) * sin2;
}
void co_zamo(double r_s, double a, double r, double theta, double& u_t, double& u_phi) {  // :59-72
    double g_tt, g_tphi, g_phiphi;
    co_metric_components_at(r_s, a, r, theta, g_tt, g_tphi, g_phiphi);
    double omega = -g_tphi / g_phiphi;
    u_t = std::sqrt(-1.0 / (g_tt + 2.0 * g_tphi * omega + g_phiphi * omega * omega));
    u_phi = omega * u_t;
}
double co_angular_velocity(double r_s, double a, double r) {  // :76-80
    double m = 0.5 * r_s;
    double sqrt_m = std::sqrt(m);
    return sqrt_m / (std::pow(r, 1.5) + a * sqrt_m);
}
// :84-107; returns false for NoCircularOrbitPossible
bool co_killing(double r_s, double a, double r, double& u_t, double& u_phi) {
    double omega = co_angular_velocity(r_s, a, r);
    double g_tt, g_tphi, g_phiphi;
    co_metric_components_at(r_s, a, r, PI / 2.0, g_tt, g_tphi, g_phiphi);
    double ut_pre = g_tt + 2.0 * omega * g_tphi + omega * omega * g_phiphi;
    if (ut_pre >= 0.0) return false;
    u_t = 1.0 / std::sqrt(-ut_pre);
    u_phi = omega * u_t;
    return true;
}
bool co_conserved_energy(double r_s, double a, double r, double& out) {  // :110-115
    double omega = co_angular_velocity(r_s, a, r);
    double g_tt, g_tphi, g_phiphi;
    co_metric_components_at(r_s, a, r, PI / 2.0, g_tt, g_tphi, g_phiphi);
    double u_t, u_phi;
    if (!co_killing(r_s, a, r, u_t, u_phi)) return false;
    out = -(g_tt + g_tphi * omega) * u_t;
    return true;
}
bool co_conserved_angular_momentum(double r_s, double a, double r, double& out) {  // :118-123
    double omega = co_angular_velocity(r_s, a, r);
    double g_tt, g_tphi, g_phiphi;
    co_metric_components_at(r_s, a, r, PI / 2.0, g_tt, g_tphi, g_phiphi);
    double u_t, u_phi;
    if (!co_killing(r_s, a, r, u_t, u_phi)) return false;
    out = (g_tphi + g_phiphi * omega) * u_t;
    return true;
}
double co_r_isco(double r_s, double a) {  // :126-136
    double a_s = 2.0 * a / r_s;
    double z1 = 1.0 + std::pow(1.0 - a_s * a_s, 1.0 / 3.0) *
                          (std::pow(1.0 + a_s, 1.0 / 3.0) + std::pow(1.0 - a_s, 1.0 / 3.0));
    double z2 = std::sqrt(3.0 * a_s * a_s + z1 * z1);
    return (3.0 + z2 - std::sqrt((3.0 - z1) * (3.0 + z1 + 2.0 * z2))) * r_s / 2.0;
}

// ---------------------------------------------------------------------------------------
// Geometry trait (inner product, radial coordinate, observers)
// ---------------------------------------------------------------------------------------
double inner_product(const Geom& g, V4 pos, V4 v, V4 w) {
    switch (g.kind) {
        case GRRT_GEOM_EUCLIDEAN:  // euclidean.rs:61-67
            return 1.0 * v[0] * w[0] + -v[1] * w[1] + -v[2] * w[2] + -v[3] * w[3];
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL: {  // euclidean_spherical.rs:80-90
            double r = pos[1], theta = pos[2];
            return 1.0 * v[0] * w[0] - v[1] * w[1] - r * r * v[2] * w[2] -
                   r * r * std::sin(theta) * std::sin(theta) * v[3] * w[3];
        }
        case GRRT_GEOM_SCHWARZSCHILD: {  // schwarzschild.rs:90-102
            double r = pos[1], theta = pos[2];
            double a = 1.0 - g.radius / r;
            return a * v[0] * w[0] - v[1] * w[1] / a - r * r * v[2] * w[2] -
                   r * r * std::sin(theta) * std::sin(theta) * v[3] * w[3];
        }
        case GRRT_GEOM_KERR: {  // kerr.rs:283-287
            M4 m = ks_metric(g.radius, g.a, pos[1], pos[2], pos[3]);
            return vt_m_w(v, m, w);
        }
        default: {  // kerr_bl.rs:339-358
            M4 m = bl_metric(g.radius, g.a, pos[1], pos[2]);
            double result = 0.0;
            for (int mu = 0; mu < 4; ++mu)
                for (int nu = 0; nu < 4; ++nu) result += m.m[mu][nu] * v[mu] * w[nu];
            return result;
        }
    }
}
// Geometry::get_radial_coordinate for a point given in chart `c`
double radial_coordinate(const Geom& g, Chart c, V4 p) {
    switch (g.kind) {
        case GRRT_GEOM_EUCLIDEAN: return get_as_spherical(c, p).x;  // euclidean.rs:157-159
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL:                            // euclidean_spherical.rs:135-146
        case GRRT_GEOM_SCHWARZSCHILD:                                  // schwarzschild.rs:201-211
            return c == CART ? get_as_spherical(c, p).x : p[1];
        case GRRT_GEOM_KERR: return std::sqrt(ks_r_sqr(g.a, p[1], p[2], p[3]));  // kerr.rs:416-419
        default:                                                                 // kerr_bl.rs:579-594
            if (c == BL) return p[1];
            {
                double x = p[1], y = p[2], z = p[3];
                double rho_sqr = x * x + y * y + z * z;
                double r_sqr = 0.5 * (rho_sqr - g.a * g.a +
                                      std::sqrt(powi2(rho_sqr - g.a * g.a) + 4.0 * g.a * g.a * z * z));
                return std::sqrt(r_sqr);
            }
    }
}
V4 axial_killing_vector(const Geom& g, V4 pos) {
    if (g.chart() == CART) return v4(0.0, -pos[2], pos[1], 0.0);  // kerr.rs:482-485, euclidean.rs:202-204
    return v4(0.0, 0.0, 0.0, 1.0);
}
V4 stationary_velocity(const Geom& g, V4 pos) {
    switch (g.kind) {
        case GRRT_GEOM_EUCLIDEAN:
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL: return v4(1.0, 0.0, 0.0, 0.0);
        case GRRT_GEOM_SCHWARZSCHILD: {  // schwarzschild.rs:237-240
            double a = 1.0 - g.radius / pos[1];
            return v4(1.0 / std::sqrt(a), 0.0, 0.0, 0.0);
        }
        case GRRT_GEOM_KERR: {  // kerr.rs:449-455
            double f = ks_f(g.radius, g.a, pos[1], pos[2], pos[3]);
            return v4(1.0 / std::sqrt(1.0 - f), 0.0, 0.0, 0.0);
        }
        default: {  // kerr_bl.rs:362-371
            double sig = bl_sigma(pos[1], g.a, pos[2]);
            double ut = 1.0 / std::sqrt(1.0 - g.radius * pos[1] / sig);
            return v4(ut, 0.0, 0.0, 0.0);
        }
    }
}
V4 zamo_velocity(const Geom& g, V4 pos) {
    switch (g.kind) {
        case GRRT_GEOM_KERR: {  // kerr.rs:457-468
            double x = pos[1], y = pos[2], z = pos[3];
            double r = std::sqrt(ks_r_sqr(g.a, x, y, z));
            double theta = (r == 0.0) ? 0.0 : std::acos(rclamp(z / r, -1.0, 1.0));
            double u_t, u_phi;
            co_zamo(g.radius, g.a, r, theta, u_t, u_phi);
            return u_t * v4(1.0, 0.0, 0.0, 0.0) + u_phi * axial_killing_vector(g, pos);
        }
        case GRRT_GEOM_KERR_BL: {  // kerr_bl.rs:373-382
            double u_t, u_phi;
            co_zamo(g.radius, g.a, pos[1], pos[2], u_t, u_phi);
            return v4(u_t, 0.0, 0.0, u_phi);
        }
        default: return stationary_velocity(g, pos);
    }
}
// circular_orbit_killing_coefficients for a point in chart `c`
bool orbit_coefficients(const Geom& g, Chart c, V4 pos, double& u_t, double& u_phi) {
    switch (g.kind) {
        case GRRT_GEOM_EUCLIDEAN:
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL:
            u_t = 1.0;
            u_phi = 0.0;
            return true;
        case GRRT_GEOM_SCHWARZSCHILD:  // schwarzschild.rs:260-265
            return co_killing(g.radius, 0.0, radial_coordinate(g, c, pos), u_t, u_phi);
        default:  // kerr.rs:487-496, kerr_bl.rs:398-410
            return co_killing(g.radius, g.a, radial_coordinate(g, c, pos), u_t, u_phi);
    }
}
// get_circular_orbit_velocity_at (native chart)
bool circular_orbit_velocity(const Geom& g, V4 pos, V4& out) {
    double u_t, u_phi;
    switch (g.kind) {
        case GRRT_GEOM_EUCLIDEAN:
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL: out = v4(1.0, 0.0, 0.0, 0.0); return true;
        case GRRT_GEOM_KERR:  // kerr.rs:473-480
            if (!orbit_coefficients(g, CART, pos, u_t, u_phi)) return false;
            out = u_t * v4(1.0, 0.0, 0.0, 0.0) + u_phi * axial_killing_vector(g, pos);
            return true;
        default:  // schwarzschild.rs:248-254, kerr_bl.rs:384-392
            if (!orbit_coefficients(g, g.chart(), pos, u_t, u_phi)) return false;
            out = v4(u_t, 0.0, 0.0, u_phi);
            return true;
    }
}

// ---------------------------------------------------------------------------------------
// tetrads, Gram-Schmidt, Lorentz boost, camera (camera.rs:27-196)
// ---------------------------------------------------------------------------------------
struct Tetrad {
    V4 t, x, y, z;
};

// gram_schmidt.rs:6-34
void gram_schmidt(const Geom& g, V4 pos, const V4 in[4], V4 out[4]) {
    for (int i = 0; i < 4; ++i) {
        V4 w = in[i];
        for (int j = 0; j < i; ++j) {
            double p1 = inner_product(g, pos, w, out[j]);
            double p2 = inner_product(g, pos, out[j], out[j]);
            V4 projection = (p1 / p2) * out[j];
            w = w + (-projection);
        }
        double n = std::sqrt(std::fabs(inner_product(g, pos, w, w)));
        out[i] = (1.0 / n) * w;
    }
}

Tetrad tetrad_at(const Geom& g, V4 pos) {
    Tetrad T;
    switch (g.kind) {
        case GRRT_GEOM_EUCLIDEAN: {  // euclidean.rs:88-110
            V3 s = get_as_spherical(CART, pos);
            double theta = s.y, phi = s.z;
            V4 e_r = v4(0.0, std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta));
            V4 e_theta = v4(0.0, std::cos(theta) * std::cos(phi), std::cos(theta) * std::sin(phi), -std::sin(theta));
            V4 e_phi = v4(0.0, -std::sin(phi), std::cos(phi), 0.0);
            T.t = v4(1.0, 0.0, 0.0, 0.0);
            T.x = e_phi;
            T.y = -e_theta;
            T.z = -e_r;
            return T;
        }
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL: {  // euclidean_spherical.rs:100-111
            double r = pos[1], theta = pos[2];
            T.t = v4(1.0, 0.0, 0.0, 0.0);
            T.x = v4(0.0, 0.0, 0.0, 1.0 / (r * std::sin(theta)));
            T.y = -v4(0.0, 0.0, 1.0 / r, 0.0);
            T.z = -v4(0.0, 1.0, 0.0, 0.0);
            return T;
        }
        case GRRT_GEOM_SCHWARZSCHILD: {  // schwarzschild.rs:116-132
            double r = pos[1], theta = pos[2];
            double rr0 = g.radius / r;
            double a = 1.0 - rr0;
            T.t = v4(1.0 / a, -std::sqrt(rr0), 0.0, 0.0);
            T.x = v4(0.0, 0.0, 0.0, 1.0 / (r * std::sin(theta)));
            T.y = v4(0.0, 0.0, 1.0 / r, 0.0);
            T.z = v4(-std::sqrt(rr0) / a, 1.0, 0.0, 0.0);
            return T;
        }
        case GRRT_GEOM_KERR: {  // kerr.rs:297-331
            double x = pos[1], y = pos[2], z = pos[3];
            double f = ks_f(g.radius, g.a, x, y, z);
            V4 k = ks_k_covector(g.a, x, y, z);
            double alpha = 1.0 / std::sqrt(1.0 + f);
            double bfac = f / (1.0 + f);
            double beta[3] = {bfac * k[1], bfac * k[2], bfac * k[3]};
            V4 in[4] = {v4(1.0 / alpha, -beta[0] / alpha, -beta[1] / alpha, -beta[2] / alpha),
                        v4(0.0, 1.0, 0.0, 0.0), v4(0.0, 0.0, 1.0, 0.0), v4(0.0, 0.0, 0.0, 1.0)};
            V4 out[4];
            gram_schmidt(g, pos, in, out);
            T.t = out[0];
            T.x = out[1];
            T.y = out[2];
            T.z = out[3];
            return T;
        }
        default: {  // kerr_bl.rs:428-450
            double r = pos[1], theta = pos[2];
            double u_t, u_phi;
            co_zamo(g.radius, g.a, r, theta, u_t, u_phi);
            double omega = u_phi / u_t, ut = u_t;  // zamo_params :72-75
            V4 in[4] = {v4(ut, 0.0, 0.0, ut * omega), v4(0.0, 0.0, 0.0, 1.0), v4(0.0, 0.0, 1.0, 0.0),
                        v4(0.0, 1.0, 0.0, 0.0)};
            V4 out[4];
            gram_schmidt(g, pos, in, out);
            T.t = out[0];
            T.x = out[1];
            T.y = out[2];
            T.z = out[3];
            return T;
        }
    }
}

M4 lorentz_transformation(const Geom& g, V4 pos, V4 vel) {
    M4 L = zeros4();
    V4 tt = tetrad_at(g, pos).t;
    switch (g.kind) {
        case GRRT_GEOM_EUCLIDEAN: {  // euclidean.rs:112-142
            double gamma = tt[0] * vel[0] - tt[1] * vel[1] - tt[2] * vel[2] - tt[3] * vel[3];
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) {
                    double res = (i == j) ? 1.0 : 0.0;
                    double gg = (j == 0) ? 1.0 : -1.0;
                    double a = -1.0 / (1.0 + gamma);
                    double b = tt[i] + vel[i];
                    double c = gg * (tt[j] + vel[j]);
                    res += a * b * c;
                    res += 2.0 * gg * tt[i] * vel[j];
                    L.m[i][j] = res;
                }
            return L;
        }
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL: {  // euclidean_spherical.rs:113-121
            for (int i = 0; i < 4; ++i) L.m[i][i] = 1.0;
            return L;
        }
        case GRRT_GEOM_SCHWARZSCHILD: {  // schwarzschild.rs:134-179
            double r = pos[1], theta = pos[2];
            double a = 1.0 - g.radius / r;
            double md[4] = {a, -1.0 / a, -r * r, -r * r * std::sin(theta) * std::sin(theta)};
            double gamma = 0.0;
            for (int i = 0; i < 4; ++i) gamma += md[i] * vel[i] * tt[i];
            for (int mu = 0; mu < 4; ++mu)
                for (int nu = 0; nu < 4; ++nu) {
                    double res = (mu == nu) ? 1.0 : 0.0;
                    double aa = 1.0 / (1.0 + gamma);
                    double b = tt[mu] + vel[mu];
                    double c = md[nu] * (tt[nu] + vel[nu]);
                    res -= aa * b * c;
                    res += 2.0 * md[nu] * tt[nu] * vel[mu];
                    L.m[mu][nu] = res;
                }
            return L;
        }
        default: {  // kerr.rs:333-380, kerr_bl.rs:452-480
            M4 m = (g.kind == GRRT_GEOM_KERR) ? ks_metric(g.radius, g.a, pos[1], pos[2], pos[3])
                                              : bl_metric(g.radius, g.a, pos[1], pos[2]);
            double gamma = -vt_m_w(tt, m, vel);
            V4 uv = tt + vel;
            V4 uv_lower = mul(m, uv);
            V4 mt = mul(m, tt);
            for (int mu = 0; mu < 4; ++mu)
                for (int nu = 0; nu < 4; ++nu) {
                    double res = (mu == nu) ? 1.0 : 0.0;
                    double a = 1.0 / (1.0 + gamma);
                    double b = uv[mu];
                    double c = uv_lower[nu];
                    res += a * b * c;
                    res -= 2.0 * mt[nu] * vel[mu];
                    L.m[mu][nu] = res;
                }
            return L;
        }
    }
}

bool tetrad_valid(const Geom& g, V4 pos, const Tetrad& T) {  // tetrad.rs:60-130
    double tol = 1e-5;
    double s0 = g.sig0(), s1 = -s0;
    auto ok = [&](double a, double b) { return std::fabs(a - b) < tol; };
    return ok(inner_product(g, pos, T.t, T.t), s0) && ok(inner_product(g, pos, T.x, T.x), s1) &&
           ok(inner_product(g, pos, T.y, T.y), s1) && ok(inner_product(g, pos, T.z, T.z), s1) &&
           ok(inner_product(g, pos, T.t, T.x), 0.0) && ok(inner_product(g, pos, T.t, T.y), 0.0) &&
           ok(inner_product(g, pos, T.t, T.z), 0.0) && ok(inner_product(g, pos, T.x, T.y), 0.0) &&
           ok(inner_product(g, pos, T.x, T.z), 0.0) && ok(inner_product(g, pos, T.y, T.z), 0.0);
}

void rotate(V4 v1, V4 v2, double angle, V4& o1, V4& o2) {  // camera.rs:76-81
    o1 = std::cos(angle) * v1 + std::sin(angle) * v2;
    o2 = -std::sin(angle) * v1 + std::cos(angle) * v2;
}

V3 spatial_basis_vector_cartesian(const Geom& g, V4 pos, V4 vec) {  // camera.rs:83-132
    Chart c = g.chart();
    if (c == CART) return v3(vec[1], vec[2], vec[3]);
    V3 s = get_as_spherical(c, pos);
    double r = s.x, theta = s.y, phi = s.z;
    double dr = vec[1], dtheta = vec[2], dphi = vec[3];
    double st = std::sin(theta), ct = std::cos(theta), sp = std::sin(phi), cp = std::cos(phi);
    if (c == SPH) {
        return v3(st * cp * dr + r * ct * cp * dtheta - r * st * sp * dphi,
                  st * sp * dr + r * ct * sp * dtheta + r * st * cp * dphi, ct * dr - r * st * dtheta);
    }
    double a = g.a;
    return v3(st * cp * dr + (r * cp - a * sp) * ct * dtheta + (-r * sp - a * cp) * st * dphi,
              st * sp * dr + (r * sp + a * cp) * ct * dtheta + (r * cp - a * sp) * st * dphi,
              ct * dr - r * st * dtheta);
}

double spatial_handedness(const Geom& g, V4 pos, const Tetrad& T) {  // camera.rs:134-148
    V3 x = spatial_basis_vector_cartesian(g, pos, T.x);
    V3 y = spatial_basis_vector_cartesian(g, pos, T.y);
    V3 z = spatial_basis_vector_cartesian(g, pos, T.z);
    double triple = dot(x, cross(y, z));
    if (!std::isfinite(triple) || std::fabs(triple) <= 1e-12) return g.chart() == CART ? 1.0 : -1.0;
    return triple >= 0.0 ? 1.0 : -1.0;
}

// Camera::new, camera.rs:151-196.  Returns 0 or GRRT_ERR_TETRAD_NOT_ORTHONORMAL.
int camera_new(const Geom& g, V4 pos, V4 vel, double alpha, int64_t rows, int64_t cols, double phi,
               double theta, double psi, GrrtCamera* out) {
    Tetrad T0 = tetrad_at(g, pos);
    if (!tetrad_valid(g, pos, T0)) return GRRT_ERR_TETRAD_NOT_ORTHONORMAL;
    V4 a_prime, b_prime, z, a_two_prime, x, y;
    rotate(T0.x, T0.y, phi, a_prime, b_prime);
    rotate(T0.z, a_prime, theta, z, a_two_prime);
    rotate(a_two_prime, b_prime, psi, x, y);
    M4 L = lorentz_transformation(g, pos, vel);  // camera.rs:27-74
    Tetrad T;
    T.t = mul(L, T0.t);
    T.x = mul(L, x);
    T.y = mul(L, y);
    T.z = mul(L, z);
    if (!tetrad_valid(g, pos, T)) return GRRT_ERR_TETRAD_NOT_ORTHONORMAL;
    std::memset(out, 0, sizeof *out);
    for (int i = 0; i < 4; ++i) {
        out->position[i] = pos[i];
        out->velocity[i] = vel[i];
        out->tetrad[0][i] = T.t[i];
        out->tetrad[1][i] = T.x[i];
        out->tetrad[2][i] = T.y[i];
        out->tetrad[3][i] = T.z[i];
    }
    out->alpha = alpha;
    out->rows = rows;
    out->cols = cols;
    out->spatial_signature = g.sig3();
    out->spatial_handedness = spatial_handedness(g, pos, T);
    return 0;
}

struct Ray {
    V4 pos, mom;
    int64_t row, col;
};

// camera.rs:214-232
V4 camera_direction(const GrrtCamera& c, double row, double column) {
    V4 et = v4(c.tetrad[0][0], c.tetrad[0][1], c.tetrad[0][2], c.tetrad[0][3]);
    (void)et;
    V4 ex = v4(c.tetrad[1][0], c.tetrad[1][1], c.tetrad[1][2], c.tetrad[1][3]);
    V4 ey = v4(c.tetrad[2][0], c.tetrad[2][1], c.tetrad[2][2], c.tetrad[2][3]);
    V4 ez = v4(c.tetrad[3][0], c.tetrad[3][1], c.tetrad[3][2], c.tetrad[3][3]);
    double shifted_column = column + 1.0;
    double shifted_row = row + 1.0;
    double tan_half_alpha = std::tan(c.alpha / 2.0);
    double i_prime = c.spatial_handedness * (2.0 * tan_half_alpha / (double)c.rows) *
                     (shifted_column - ((double)c.cols + 1.0) / 2.0);
    double j_prime = (2.0 * tan_half_alpha / (double)c.rows) * (shifted_row - ((double)c.rows + 1.0) / 2.0);
    V4 w = ez + i_prime * ex + j_prime * ey;
    double w_squared = c.spatial_signature * (1.0 + i_prime * i_prime + j_prime * j_prime);
    return -ez + (2.0 * w) / (c.spatial_signature * w_squared);
}
Ray camera_ray(const GrrtCamera& c, int64_t row, int64_t col) {  // camera.rs:234-245
    V4 d = camera_direction(c, (double)row, (double)col);
    V4 et = v4(c.tetrad[0][0], c.tetrad[0][1], c.tetrad[0][2], c.tetrad[0][3]);
    Ray r;
    r.pos = v4(c.position[0], c.position[1], c.position[2], c.position[3]);
    r.mom = d + (-et);
    r.row = row;
    r.col = col;
    return r;
}
Ray camera_ray_offset(const GrrtCamera& c, int64_t row, int64_t col, double dx, double dy) {  // :247-254
    V4 d = camera_direction(c, (double)row + (dy - 0.5), (double)col + (dx - 0.5));
    V4 et = v4(c.tetrad[0][0], c.tetrad[0][1], c.tetrad[0][2], c.tetrad[0][3]);
    Ray r;
    r.pos = v4(c.position[0], c.position[1], c.position[2], c.position[3]);
    r.mom = d + (-et);
    r.row = row;
    r.col = col;
    return r;
}

// ---------------------------------------------------------------------------------------
// geodesic solvers (GeodesicSolver trait: geometry.rs:15-32)
// ---------------------------------------------------------------------------------------
struct Solver {
    const Geom* g;
    double e, l_z, q;  // KerrBL per-ray constants
};

// kerr_bl.rs:505-577 (BL-chart branch: rays are always created in the native chart)
Solver make_solver(const Geom& g, const Ray& ray) {
    Solver s;
    s.g = &g;
    s.e = s.l_z = s.q = 0.0;
    if (g.kind == GRRT_GEOM_KERR_BL) {
        double r = ray.pos[1], theta = ray.pos[2];
        M4 m = bl_metric(g.radius, g.a, r, theta);
        V4 p_cov = mul(m, ray.mom);
        double e = -p_cov[0];
        double l_z = p_cov[3];
        double p_theta = p_cov[2];
        double cos_t = std::cos(theta), sin_t = std::sin(theta);
        double sin2 = sin_t * sin_t;
        double q = p_theta * p_theta + cos_t * cos_t * (l_z * l_z / rmax(sin2, SIN2_FLOOR) - g.a * g.a * e * e);
        s.e = e;
        s.l_z = l_z;
        s.q = q;
    }
    return s;
}

Y8 geodesic(const Solver& s, const Y8& y) {
    const Geom& g = *s.g;
    Y8 d;
    switch (g.kind) {
        case GRRT_GEOM_EUCLIDEAN:  // euclidean.rs:46-53
            d = Y8{{y[4], y[5], y[6], y[7], 0.0, 0.0, 0.0, 0.0}};
            return d;
        case GRRT_GEOM_EUCLIDEAN_SPHERICAL: {  // euclidean_spherical.rs:46-70
            double r = y[1], theta = y[2];
            double v_t = y[4], v_r = y[5], v_theta = y[6], v_phi = y[7];
            double a_t = 0.0;
            double a_r = r * (v_theta * v_theta + v_phi * v_phi * std::sin(theta) * std::sin(theta));
            double a_theta = -(2.0 / r) * v_r * v_theta + std::sin(theta) * std::cos(theta) * v_phi * v_phi;
            double a_phi = -(2.0 / r) * v_phi * v_r - 2.0 * std::cos(theta) / std::sin(theta) * v_theta * v_phi;
            d = Y8{{v_t, v_r, v_theta, v_phi, a_t, a_r, a_theta, a_phi}};
            return d;
        }
        case GRRT_GEOM_SCHWARZSCHILD: {  // schwarzschild.rs:54-80
            double r = y[1], theta = y[2];
            double v_t = y[4], v_r = y[5], v_theta = y[6], v_phi = y[7];
            double a = 1.0 - g.radius / r;
            double a_prime = g.radius / (r * r);
            double aprime_over_a = a_prime / a;
            double a_t = -(aprime_over_a)*v_t * v_r;
            double a_r = -0.5 * a * a_prime * v_t * v_t + 0.5 * (aprime_over_a)*v_r * v_r +
                         a * r * (v_theta * v_theta + v_phi * v_phi * std::sin(theta) * std::sin(theta));
            double a_theta = -(2.0 / r) * v_r * v_theta + std::sin(theta) * std::cos(theta) * v_phi * v_phi;
            double a_phi = -(2.0 / r) * v_phi * v_r - 2.0 * std::cos(theta) / std::sin(theta) * v_theta * v_phi;
            d = Y8{{v_t, v_r, v_theta, v_phi, a_t, a_r, a_theta, a_phi}};
            return d;
        }
        case GRRT_GEOM_KERR: {  // kerr.rs:200-241
            double x = y[1], yy = y[2], z = y[3];
            V4 p = v4(y[4], y[5], y[6], y[7]);
            M4 gc = ks_metric_contravariant(g.radius, g.a, x, yy, z);
            V4 xdot = mul(gc, p);
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            {
                for (int idx = 1; idx <= 3; ++idx) {
                    M4 dcov = ks_d_metric_cov_fd(g.radius, g.a, idx, x, yy, z);
                    M4 t = mul(mul(gc, dcov), gc);  // kerr.rs:149-159
                    for (int i = 0; i < 4; ++i)
                        for (int j = 0; j < 4; ++j) t.m[i][j] = -t.m[i][j];
                    acc[idx] = -0.5 * vt_m_w(p, t, p);
                }
            }
            d = Y8{{xdot[0], xdot[1], xdot[2], xdot[3], 0.0, acc[1], acc[2], acc[3]}};
            return d;
        }
        default: {  // kerr_bl.rs:141-174
            double r = y[1], theta = y[2], v_r = y[4], v_theta = y[5];
            double a = g.a;
            double del = bl_delta(r, g.radius, a);
            double p_r = (r * r + a * a) * s.e - a * s.l_z;
            double sin_t = std::sin(theta);
            double sin2 = sin_t * sin_t;
            double dt = (r * r + a * a) / del * p_r + a * (s.l_z - a * s.e * sin2);
            double dphi = a / del * p_r + s.l_z / sin2 - a * s.e;
            double dv_r = bl_potential_r_derivative(r, g.radius, a, s.e, s.l_z, s.q) / 2.0;
            double dv_theta = bl_potential_theta_derivative(theta, a, s.e, s.l_z) / 2.0;
            d = Y8{{dt, v_r, v_theta, dphi, dv_r, dv_theta, 0.0, 0.0}};
            return d;
        }
    }
}

Y8 create_initial_state(const Solver& s, const Ray& ray) {
    const Geom& g = *s.g;
    if (g.kind == GRRT_GEOM_KERR) {  // kerr.rs:243-260
        M4 m = ks_metric(g.radius, g.a, ray.pos[1], ray.pos[2], ray.pos[3]);
        V4 pc = mul(m, ray.mom);
        return Y8{{ray.pos[0], ray.pos[1], ray.pos[2], ray.pos[3], pc[0], pc[1], pc[2], pc[3]}};
    }
    if (g.kind == GRRT_GEOM_KERR_BL) {  // kerr_bl.rs:176-223 (BL branch)
        double r = ray.pos[1], theta = ray.pos[2], phi = ray.pos[3], t = ray.pos[0];
        double sign_r = ray.mom[1] >= 0.0 ? 1.0 : -1.0;
        double sign_theta = ray.mom[2] >= 0.0 ? 1.0 : -1.0;
        double r_pot = bl_potential_r(r, g.radius, g.a, s.e, s.l_z, s.q);
        double th_pot = bl_potential_theta(theta, g.a, s.e, s.l_z, s.q);
        double v_r = sign_r * std::sqrt(rmax(r_pot, 0.0));
        double v_theta = sign_theta * std::sqrt(rmax(th_pot, 0.0));
        return Y8{{t, r, theta, phi, v_r, v_theta, 0.0, 0.0}};
    }
    return Y8{{ray.pos[0], ray.pos[1], ray.pos[2], ray.pos[3], ray.mom[0], ray.mom[1], ray.mom[2], ray.mom[3]}};
}

V4 momentum_from_state(const Solver& s, const Y8& y) {
    const Geom& g = *s.g;
    if (g.kind == GRRT_GEOM_KERR) {  // kerr.rs:262-273
        M4 gc = ks_metric_contravariant(g.radius, g.a, y[1], y[2], y[3]);
        return mul(gc, v4(y[4], y[5], y[6], y[7]));
    }
    if (g.kind == GRRT_GEOM_KERR_BL) {  // kerr_bl.rs:225-249
        double r = y[1], theta = y[2], v_r = y[4], v_theta = y[5];
        double a = g.a;
        double del = bl_delta(r, g.radius, a);
        double sig = bl_sigma(r, a, theta);
        double sin2 = powi2(std::sin(theta));
        double p_r_term = (r * r + a * a) * s.e - a * s.l_z;
        double dt_dlambda = (r * r + a * a) / del * p_r_term + a * (s.l_z - a * s.e * sin2);
        double dphi_dlambda = a / del * p_r_term + s.l_z / sin2 - a * s.e;
        return v4(dt_dlambda / sig, v_r / sig, v_theta / sig, dphi_dlambda / sig);
    }
    return v4(y[4], y[5], y[6], y[7]);
}

bool inside_horizon(const Geom& g, const Y8& y) {
    switch (g.kind) {
        case GRRT_GEOM_SCHWARZSCHILD: return y[1] <= g.radius + g.heps;  // schwarzschild.rs:181-183
        case GRRT_GEOM_KERR: {  // kerr.rs:382-394
            if (std::fabs(g.a) > g.radius / 2.0) return false;
            double r = std::sqrt(ks_r_sqr(g.a, y[1], y[2], y[3]));
            double m = 0.5 * g.radius;
            double disc = rmax(m * m - g.a * g.a, 0.0);
            double rp = m + std::sqrt(disc);
            return r <= rp + g.heps;
        }
        case GRRT_GEOM_KERR_BL: {  // kerr_bl.rs:482-492
            double m = g.radius / 2.0;
            if (std::fabs(g.a) > m) return false;
            double disc = rmax(m * m - g.a * g.a, 0.0);
            double r_plus = m + std::sqrt(disc);
            return y[1] <= r_plus + g.heps;
        }
        default: return false;
    }
}
bool closed_orbit(const Geom& g, const Y8& y, uint64_t step_index, uint64_t max_steps) {
    const double F = 5.0;  // geometry.rs:106 TRAPPED_ORBIT_RADIUS_FACTOR
    switch (g.kind) {
        case GRRT_GEOM_SCHWARZSCHILD:  // schwarzschild.rs:185-193
        case GRRT_GEOM_KERR_BL:        // kerr_bl.rs:494-503
            return step_index == max_steps - 1 && y[1] < F * g.radius;
        case GRRT_GEOM_KERR: {  // kerr.rs:396-407
            double r = std::sqrt(ks_r_sqr(g.a, y[1], y[2], y[3]));
            return step_index == max_steps - 1 && r < F * g.radius;
        }
        default: return false;
    }
}

// ---------------------------------------------------------------------------------------
// RKF45 (rendering/runge_kutta.rs)
// ---------------------------------------------------------------------------------------
const double A2 = 2.0 / 9.0, A3 = 1.0 / 3.0, A4 = 3.0 / 4.0, A5 = 1.0, A6 = 5.0 / 6.0;
const double B_2_1 = 2.0 / 9.0, B_3_1 = 1.0 / 12.0, B_4_1 = 69.0 / 128.0, B_5_1 = -17.0 / 12.0, B_6_1 = 65.0 / 432.0;
const double B_3_2 = 1.0 / 4.0, B_4_2 = -243.0 / 128.0, B_5_2 = 27.0 / 4.0, B_6_2 = -5.0 / 16.0;
const double B_4_3 = 135.0 / 64.0, B_5_3 = -27.0 / 5.0, B_6_3 = 13.0 / 16.0;
const double B_5_4 = 16.0 / 15.0, B_6_4 = 4.0 / 27.0;
const double B_6_5 = 5.0 / 144.0;
const double CH_1 = 47.0 / 450.0, CH_2 = 0.0, CH_3 = 12.0 / 25.0, CH_4 = 32.0 / 225.0, CH_5 = 1.0 / 30.0, CH_6 = 6.0 / 25.0;
const double CT_1 = 1.0 / 150.0, CT_2 = 0.0, CT_3 = -3.0 / 100.0, CT_4 = 16.0 / 75.0, CT_5 = 1.0 / 20.0, CT_6 = -6.0 / 25.0;
const double BETA = 0.9, CONVERGENCY_ORDER = 5.0, ERROR_RATIO_SMALL_ERROR = 1e-5;
const int MAX_RETRY_STEP = 100;
const double H_MAX = 1.0, H_MIN = 1e-12, H_GROWTH_CAP = 4.0;

inline Y8 operator+(const Y8& a, const Y8& b) {
    Y8 r;
    for (int i = 0; i < 8; ++i) r[i] = a[i] + b[i];
    return r;
}
inline Y8 operator*(double s, const Y8& a) {
    Y8 r;
    for (int i = 0; i < 8; ++i) r[i] = s * a[i];
    return r;
}
double norm8(const Y8& a) {  // nalgebra dotc, 8 accumulators
    double q[8];
    for (int i = 0; i < 8; ++i) q[i] = a[i] * a[i];
    double res = 0.0;
    res += q[0] + q[4];
    res += q[1] + q[5];
    res += q[2] + q[6];
    res += q[3] + q[7];
    return std::sqrt(res);
}

template <class F>
void rkf45_step(const Y8& y, double h, F f, Y8& y_new, double& truncation_error) {  // :86-125
    Y8 k1 = h * f(y);
    Y8 k2 = h * f(y + B_2_1 * k1);
    Y8 k3 = h * f(y + B_3_1 * k1 + B_3_2 * k2);
    Y8 k4 = h * f(y + B_4_1 * k1 + B_4_2 * k2 + B_4_3 * k3);
    Y8 k5 = h * f(y + B_5_1 * k1 + B_5_2 * k2 + B_5_3 * k3 + B_5_4 * k4);
    Y8 k6 = h * f(y + B_6_1 * k1 + B_6_2 * k2 + B_6_3 * k3 + B_6_4 * k4 + B_6_5 * k5);
    y_new = y + CH_1 * k1 + CH_2 * k2 + CH_3 * k3 + CH_4 * k4 + CH_5 * k5 + CH_6 * k6;
    truncation_error = norm8(CT_1 * k1 + CT_2 * k2 + CT_3 * k3 + CT_4 * k4 + CT_5 * k5 + CT_6 * k6);
}

// runge_kutta.rs:138-182; returns false for Err(MaxStepsReached)
template <class F>
bool rkf45(const Y8& y, double h, double epsilon, F f, Y8& y_new, double& h_taken, double& h_next,
           uint64_t& attempts) {
    (void)A2; (void)A3; (void)A4; (void)A5; (void)A6;  // the RHS never reads t
    double h_cur = rclamp(h, H_MIN, H_MAX);
    for (int i = 0; i < MAX_RETRY_STEP; ++i) {
        double truncation_error;
        rkf45_step(y, h_cur, f, y_new, truncation_error);
        ++attempts;
        double h_proposed;
        if (truncation_error > 0.0)
            h_proposed = BETA * h_cur * std::pow(epsilon / truncation_error, 1.0 / CONVERGENCY_ORDER);
        else
            h_proposed = h_cur * H_GROWTH_CAP;
        h_proposed = rclamp(rmin(h_proposed, h_cur * H_GROWTH_CAP), H_MIN, H_MAX);
        if (truncation_error > epsilon) {
            if (h_cur <= H_MIN) {
                h_taken = h_cur;
                h_next = h_cur;
                return true;
            }
            h_cur = rclamp(h_proposed / 2.0, H_MIN, H_MAX);
        } else {
            h_next = (truncation_error / epsilon < ERROR_RATIO_SMALL_ERROR)
                         ? rclamp(h_cur * H_GROWTH_CAP, H_MIN, H_MAX)
                         : h_proposed;
            h_taken = h_cur;
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// colour (rendering/color.rs)
// ---------------------------------------------------------------------------------------
struct Cie {
    double x, y, z, alpha;
};
inline Cie cie(double x, double y, double z, double a) { return Cie{x, y, z, a}; }
Cie blend(Cie self, Cie other) {  // color.rs:49-69
    double ab = rclamp(self.alpha, 0.0, 1.0);
    double af = rclamp(other.alpha, 0.0, 1.0);
    double ao = af + ab * (1.0 - af);
    if (ao <= 0.0) return cie(0.0, 0.0, 0.0, 0.0);
    double x = (other.x * af + self.x * ab * (1.0 - af)) / ao;
    double y = (other.y * af + self.y * ab * (1.0 - af)) / ao;
    double z = (other.z * af + self.z * ab * (1.0 - af)) / ao;
    return cie(x, y, z, ao);
}
Cie apply_beaming(Cie c, double redshift, double beaming_exponent) {  // color.rs:72-80
    double f = std::pow(redshift, beaming_exponent);
    return cie(c.x * f, c.y * f, c.z * f, c.alpha);
}
double inv_compand_srgb(double u) {  // color.rs:301-308
    if (u <= 0.04045) return u / 12.92;
    return std::pow((u + 0.055) / 1.055, 2.4);
}
Cie srgb_to_xyz(uint8_t cr, uint8_t cg, uint8_t cb) {  // color.rs:310-332
    double r = inv_compand_srgb((double)cr / 255.0);
    double g = inv_compand_srgb((double)cg / 255.0);
    double b = inv_compand_srgb((double)cb / 255.0);
    double X = 0.4124564 * r + 0.3575761 * g + 0.1804375 * b;
    double Y = 0.2126729 * r + 0.7151522 * g + 0.0721750 * b;
    double Z = 0.0193339 * r + 0.1191920 * g + 0.9503041 * b;
    return cie(X, Y, Z, 1.0);
}
Cie cie_from_rgba(const uint8_t* p) {  // color.rs:32-40
    Cie c = srgb_to_xyz(p[0], p[1], p[2]);
    c.alpha = (double)p[3] / 255.0;
    return c;
}
double compand_srgb(double linear) {  // color.rs:193-202
    double sign = linear < 0.0 ? -1.0 : 1.0;
    double a = std::fabs(linear);
    double encoded = (a <= 0.0031308) ? 12.92 * a : 1.055 * std::pow(a, 1.0 / 2.4) - 0.055;
    return rclamp(sign * encoded, 0.0, 1.0);
}
void xyz_to_linear_srgb(const double* xyz, double out[3]) {  // color.rs:204-218
    out[0] = 3.2406255 * xyz[0] + -1.5372080 * xyz[1] + -0.4986286 * xyz[2];
    out[1] = -0.9689307 * xyz[0] + 1.8757561 * xyz[1] + 0.0415175 * xyz[2];
    out[2] = 0.0557101 * xyz[0] + -0.2040211 * xyz[1] + 1.0569959 * xyz[2];
}
uint8_t to_u8(double v) {  // `(x*255).round() as u8`
    double r = std::round(v);
    if (!(r > 0.0)) return 0;
    if (r >= 255.0) return 255;
    return (uint8_t)r;
}
// colour matching function fits, color.rs:173-190
double cmf_g(double lambda, double mu, double tau_left, double tau_right) {
    double tau = lambda < mu ? tau_left : tau_right;
    double t = (lambda - mu) * tau;
    return std::exp(-0.5 * t * t);
}
double x_bar(double l) {
    return 1.056 * cmf_g(l, 599.8, 0.0264, 0.0323) + 0.362 * cmf_g(l, 442.0, 0.0624, 0.0374) -
           0.065 * cmf_g(l, 501.1, 0.0490, 0.0382);
}
double y_bar(double l) { return 0.821 * cmf_g(l, 568.8, 0.0213, 0.0247) + 0.286 * cmf_g(l, 530.9, 0.0613, 0.0322); }
double z_bar(double l) { return 1.217 * cmf_g(l, 437.0, 0.0845, 0.0278) + 0.681 * cmf_g(l, 459.0, 0.0385, 0.0725); }

// black_body_radiation.rs:12-41
double planck_spectral_radiance(double lambda, double temperature) {
    const double H = 6.62607015e-34, C = 299792458.0, KB = 1.380649e-23;
    double a = 2.0 * H * C * C;
    double b = H * C / (lambda * KB * temperature);
    return a / (powi5(lambda) * (std::exp(b) - 1.0));
}
Cie integrate_blackbody_xyz(double temperature, double redshift) {
    const double NM = 1e-9;
    double interval = (830.0 - 380.0) * NM;
    double step_size = 1.0 * NM;
    uint64_t num_steps = sat_u64(std::floor(interval / step_size));
    double xa = 0.0, ya = 0.0, za = 0.0;
    for (uint64_t i = 0; i < num_steps; ++i) {
        double lambda = 380.0 * NM + ((double)i + 0.5) * step_size;
        double radiance = planck_spectral_radiance(lambda * redshift, temperature);
        xa += radiance * x_bar(lambda / NM) * step_size;
        ya += radiance * y_bar(lambda / NM) * step_size;
        za += radiance * z_bar(lambda / NM) * step_size;
    }
    double boost = powi5(redshift);
    return cie(xa * boost, ya * boost, za * boost, 1.0);
}

// ---------------------------------------------------------------------------------------
// textures (rendering/texture.rs)
// ---------------------------------------------------------------------------------------
struct SceneView {  // borrowed view of a GrrtScene
    const GrrtScene* s;
    Geom g;
};

Cie bitmap_bilinear(const GrrtTexture& t, double u, double v) {  // texture.rs:62-90
    uint32_t width = t.width, height = t.height;
    double p_x = (double)width * u;
    double p_y = (double)height * v;
    uint32_t xf = std::min(sat_u32(std::floor(p_x)), width - 1);
    uint32_t yf = std::min(sat_u32(std::floor(p_y)), height - 1);
    uint32_t xc = std::min(sat_u32(std::ceil(p_x)), width - 1);
    uint32_t yc = std::min(sat_u32(std::ceil(p_y)), height - 1);
    auto px = [&](uint32_t x, uint32_t y) { return cie_from_rgba(t.rgba + 4 * ((size_t)y * width + x)); };
    Cie c00 = px(xf, yf), c01 = px(xf, yc), c11 = px(xc, yc), c10 = px(xc, yf);
    double dx = p_x - (double)xf;
    double dy = p_y - (double)yf;
    double w00 = (1.0 - dx) * (1.0 - dy), w01 = (1.0 - dx) * dy, w10 = dx * (1.0 - dy), w11 = dx * dy;
    auto sc = [](double w, Cie c) { return cie(w * c.x, w * c.y, w * c.z, w * c.alpha); };
    auto ad = [](Cie a, Cie b) { return cie(a.x + b.x, a.y + b.y, a.z + b.z, a.alpha + b.alpha); };
    return ad(ad(ad(sc(w00, c00), sc(w10, c10)), sc(w01, c01)), sc(w11, c11));
}
Cie sample_blackbody(const GrrtScene& s, double temperature) {  // texture.rs:149-194
    uint32_t n = s.bb_lut_n;
    if (n == 0) return cie(0.0, 0.0, 0.0, 1.0);
    const double* lt = s.bb_lut_log_t;
    const double* c = s.bb_lut_xyz;
    auto entry = [&](uint32_t i) { return cie(c[3 * i], c[3 * i + 1], c[3 * i + 2], 1.0); };
    double log_t = std::log10(rmax(temperature, 10.0));
    if (!std::isfinite(log_t)) return entry(0);
    if (log_t <= lt[0]) return entry(0);
    if (log_t >= lt[n - 1]) return entry(n - 1);
    // binary_search_by(total_cmp): Ok(i) => i, Err(i) => i-1  ==  last key <= log_t
    uint32_t idx = (uint32_t)(std::upper_bound(lt, lt + n, log_t) - lt) - 1;
    double lt0 = lt[idx], lt1 = lt[idx + 1];
    Cie c0 = entry(idx), c1 = entry(idx + 1);
    double t = (log_t - lt0) / (lt1 - lt0);
    return cie(c0.x + t * (c1.x - c0.x), c0.y + t * (c1.y - c0.y), c0.z + t * (c1.z - c0.z), 1.0);
}
Cie texture_color(const GrrtScene& s, const GrrtTexture& t, double u, double v, double temperature,
                  double redshift) {
    switch (t.kind) {
        case GRRT_TEX_BITMAP:  // texture.rs:93-102
            return apply_beaming(bitmap_bilinear(t, u, v), redshift, t.beaming_exponent);
        case GRRT_TEX_CHECKER: {  // texture.rs:240-256
            uint64_t ut = sat_u64(std::floor(u * t.checker_width));
            uint64_t vt = sat_u64(std::floor(v * t.checker_height));
            const double* c = ((ut + vt) % 2 == 0) ? t.c1 : t.c2;
            return apply_beaming(cie(c[0], c[1], c[2], c[3]), redshift, t.beaming_exponent);
        }
        default:  // texture.rs:198-210: LUT at T*z, then the artistic exponent
            return apply_beaming(sample_blackbody(s, temperature * redshift), redshift, t.beaming_exponent);
    }
}

// temperature.rs:197-253; returns GRRT_RAYERR_* (0 = ok)
int compute_temperature(const GrrtTemperature& tc, double radius, double& out) {
    if (tc.kind == GRRT_TEMP_CONSTANT) {
        out = tc.constant;
        return 0;
    }
    if (!std::isfinite(radius)) return GRRT_RAYERR_NONFINITE_RADIUS;
    if (radius < tc.r_isco) return GRRT_RAYERR_BELOW_RISCO;
    if (tc.n == 0) return GRRT_RAYERR_BELOW_RISCO;
    const double* rr = tc.lut_r;
    const double* tt = tc.lut_t;
    uint32_t n = tc.n;
    if (radius <= rr[0]) {
        out = tt[0];
        return 0;
    }
    if (radius >= rr[n - 1]) {
        out = tt[n - 1];
        return 0;
    }
    uint32_t idx = (uint32_t)(std::upper_bound(rr, rr + n, radius) - rr) - 1;
    double r0 = rr[idx], t0 = tt[idx], r1 = rr[idx + 1], t1 = tt[idx + 1];
    double t = (radius - r0) / (r1 - r0);
    out = t0 + t * (t1 - t0);
    return 0;
}

// ---------------------------------------------------------------------------------------
// Perlin noise: crate noise 0.9.0 (core/perlin.rs perlin_3d, permutationtable.rs), with
// rand_xorshift 0.3 XorShiftRng and rand 0.8 SliceRandom::shuffle.  UNPINNED (see header).
// ---------------------------------------------------------------------------------------
struct XorShift {
    uint32_t x, y, z, w;
    uint32_t next() {
        uint32_t t = x ^ (x << 11);
        x = y;
        y = z;
        z = w;
        w = w ^ (w >> 19) ^ (t ^ (t >> 8));
        return w;
    }
};
uint32_t gen_range_u32(XorShift& rng, uint32_t range) {  // rand 0.8 UniformInt::sample_single, low = 0
    int lz = __builtin_clz(range);
    uint32_t zone = (range << lz) - 1u;
    for (;;) {
        uint32_t v = rng.next();
        uint64_t m = (uint64_t)v * (uint64_t)range;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo <= zone) return hi;
    }
}
void perlin_permutation(uint32_t seed, uint8_t out[256]) {
    uint8_t real[16] = {0};
    real[0] = 1;
    for (int i = 1; i < 4; ++i) {
        real[i * 4] = (uint8_t)seed;
        real[i * 4 + 1] = (uint8_t)(seed >> 8);
        real[i * 4 + 2] = (uint8_t)(seed >> 16);
        real[i * 4 + 3] = (uint8_t)(seed >> 24);
    }
    uint32_t s[4];
    for (int i = 0; i < 4; ++i)
        s[i] = (uint32_t)real[4 * i] | ((uint32_t)real[4 * i + 1] << 8) | ((uint32_t)real[4 * i + 2] << 16) |
               ((uint32_t)real[4 * i + 3] << 24);
    if (!s[0] && !s[1] && !s[2] && !s[3]) s[0] = s[1] = s[2] = s[3] = 0x0BAD5EEDu;
    XorShift rng{s[0], s[1], s[2], s[3]};
    for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
    for (uint32_t i = 255; i >= 1; --i) {  // shuffle: for i in (1..len).rev() swap(i, gen_index(i+1))
        uint32_t j = gen_range_u32(rng, i + 1);
        std::swap(out[i], out[j]);
    }
}
inline unsigned perlin_hash(const uint8_t* perm, int64_t x, int64_t y, int64_t z) {
    unsigned a = (unsigned)(x & 0xff), b = (unsigned)(y & 0xff), c = (unsigned)(z & 0xff);
    unsigned idx = perm[a] ^ b;
    idx = perm[idx] ^ c;
    return perm[idx];
}
inline double perlin_grad(unsigned h, double x, double y, double z) {
    switch (h & 15u) {
        case 0: case 12: return x + y;
        case 1: case 13: return -x + y;
        case 2: return x - y;
        case 3: return -x - y;
        case 4: return x + z;
        case 5: return -x + z;
        case 6: return x - z;
        case 7: return -x - z;
        case 8: return y + z;
        case 9: case 14: return -y + z;
        case 10: return y - z;
        default: return -y - z;  // 11 | 15
    }
}
inline double quintic(double x) { return x * x * x * (x * (x * 6.0 - 15.0) + 10.0); }
double perlin_3d(const uint8_t* perm, double px, double py, double pz) {
    const double SCALE_FACTOR = 1.1547005383792515;  // 2/sqrt(3)
    double fx = std::floor(px), fy = std::floor(py), fz = std::floor(pz);
    int64_t cx = (int64_t)fx, cy = (int64_t)fy, cz = (int64_t)fz;
    double dx = px - fx, dy = py - fy, dz = pz - fz;
    auto gr = [&](int ox, int oy, int oz) {
        return perlin_grad(perlin_hash(perm, cx + ox, cy + oy, cz + oz), dx - (double)ox, dy - (double)oy,
                           dz - (double)oz);
    };
    double g000 = gr(0, 0, 0), g100 = gr(1, 0, 0), g010 = gr(0, 1, 0), g110 = gr(1, 1, 0);
    double g001 = gr(0, 0, 1), g101 = gr(1, 0, 1), g011 = gr(0, 1, 1), g111 = gr(1, 1, 1);
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
    return rclamp(unscaled * SCALE_FACTOR, -1.0, 1.0);
}

// ---------------------------------------------------------------------------------------
// scene objects (scene_objects/*.rs)
// ---------------------------------------------------------------------------------------
struct Intersection {
    double u, v;
    V4 point;  // Cartesian, t = 0
    V3 direction;
    double t;
};
struct Frequency {  // redshift.rs:15-20
    double observer_energy, p_t, p_phi;
};
struct StepRec {  // integrator.rs:13-19
    V4 x, p;
    double t;
    uint64_t step;
};

// disc.rs:41-88
bool disc_intersects(const GrrtObject& o, V3 ys, V3 ye, Intersection& out) {
    V3 normal = v3(0.0, 0.0, 1.0), center = v3(0.0, 0.0, 0.0);
    V3 direction = ye - ys;
    double p1 = dot(center - ys, normal);
    double p2 = dot(direction, normal);
    double t = p1 / p2;
    if (!in01(t)) return false;
    V3 ip = ys + t * direction;
    double rr = norm_sq(ip);
    if (rr >= o.inner_radius * o.inner_radius && rr <= o.outer_radius * o.outer_radius) {
        V3 vp = ip - center;
        double phi = std::atan2(vp.y, vp.x);
        double r = (std::sqrt(rr) - o.inner_radius) / (o.outer_radius - o.inner_radius);
        out.u = 0.5 + 0.5 * r * std::cos(phi);
        out.v = 0.5 + 0.5 * r * std::sin(phi);
        out.point = v4(0.0, ip.x, ip.y, ip.z);
        out.t = t;
        out.direction = direction;
        return true;
    }
    return false;
}

// sphere.rs:37-57
bool sphere_solve_for_t(V3 ys, V3 d, double r, double& t) {
    double a = dot(d, d);
    double b = 2.0 * dot(ys, d);
    double c = dot(ys, ys) - r * r;
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return false;
    double sq = std::sqrt(disc);
    double t1 = (-b + sq) / (2.0 * a);
    double t2 = (-b - sq) / (2.0 * a);
    if (in01(t1)) {
        t = t1;
        return true;
    }
    if (in01(t2)) {
        t = t2;
        return true;
    }
    return false;
}
// sphere.rs:61-127 (inputs already converted by to_cartesian)
bool sphere_intersects(const GrrtObject& o, V3 ys_c, V3 ye_c, Intersection& out) {
    V3 pos = v3(o.position[0], o.position[1], o.position[2]);
    V3 neg = v3(-pos.x, -pos.y, -pos.z);
    V3 ys = ys_c + neg, ye = ye_c + neg;
    double r_start = ys.x * ys.x + ys.y * ys.y + ys.z * ys.z;
    double r_end = ye.x * ye.x + ye.y * ye.y + ye.z * ye.z;
    double R2 = o.radius * o.radius;
    if ((r_start >= R2 && r_end <= R2) || (r_start <= R2 && r_end >= R2)) {
        V3 direction = ye - ys;
        double t;
        if (!sphere_solve_for_t(ys, direction, o.radius, t)) return false;
        V3 ps = ys + t * direction;
        V4 local = cartesian_to_spherical(v4(0.0, ps.x, ps.y, ps.z));
        double theta = local[2], phi = local[3];
        double u = (PI + phi) / (2.0 * PI);
        double v = theta / PI;
        out.u = 1.0 - u;
        out.v = v;
        out.point = v4(0.0, ps.x + pos.x, ps.y + pos.y, ps.z + pos.z);
        out.t = t;
        out.direction = direction;
        return true;
    }
    return false;
}

// --- VolumetricDisc (volumetric_disc.rs) ---
const double MIN_INTERSECTION_T = 1e-9;
enum CylKind { CYL_NONE, CYL_PARALLEL, CYL_ONE, CYL_TWO };
struct CylHit {
    CylKind kind;
    double t1, t2;
};
inline V3 axis_of(const GrrtObject& o) { return v3(o.axis[0], o.axis[1], o.axis[2]); }

CylHit vol_clipped_cylinder(V3 from, V3 to, double radius, V3 axis, double half_height) {  // :348-405
    V3 seg = to - from;
    double seg_len = norm(seg);
    if (seg_len < 1e-12) return CylHit{CYL_NONE, 0, 0};
    V3 d = seg / seg_len;
    V3 v = cross(from, axis);
    V3 w = cross(d, axis);
    double a = dot(w, w);
    double b = 2.0 * dot(v, w);
    double c = dot(v, v) - radius * radius;
    if (a < 1e-10) {
        if (norm_sq(v) > radius * radius) return CylHit{CYL_NONE, 0, 0};
        return CylHit{CYL_PARALLEL, 0, 0};
    }
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return CylHit{CYL_NONE, 0, 0};
    double sq = std::sqrt(disc);
    double dists[2] = {(-b - sq) / (2.0 * a), (-b + sq) / (2.0 * a)};
    double hits[2];
    int n = 0;
    for (int k = 0; k < 2; ++k) {
        double t = dists[k] / seg_len;
        if (in01(t)) {
            V3 p = from + t * (to - from);
            if (std::fabs(dot(p, axis)) <= half_height) hits[n++] = t;
        }
    }
    if (n == 0) return CylHit{CYL_NONE, 0, 0};
    if (n == 1) return CylHit{CYL_ONE, hits[0], 0};
    return CylHit{CYL_TWO, rmin(hits[0], hits[1]), rmax(hits[0], hits[1])};
}
bool vol_cap(V3 from, V3 to, double radius, V3 axis, double pos, double& t_out) {  // :407-440
    V3 seg = to - from;
    double seg_len = norm(seg);
    if (seg_len < 1e-12) return false;
    V3 n = seg / seg_len;
    double denom = dot(n, axis);
    if (std::fabs(denom) < 1e-10) return false;
    double t = (pos - dot(from, axis)) / dot(seg, axis);
    if (!in01(t)) return false;
    V3 p = from + t * seg;
    if (norm_sq(cross(p, axis)) > radius * radius) return false;
    t_out = t;
    return true;
}
static bool total_less(double a, double b) {  // f64::total_cmp ordering
    int64_t x, y;
    std::memcpy(&x, &a, 8);
    std::memcpy(&y, &b, 8);
    x ^= (int64_t)((uint64_t)(x >> 63) >> 1);
    y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
    return x < y;
}
CylHit vol_cylinder(const GrrtObject& o, V3 from, V3 to) {  // :442-494
    double hits[6];
    int n = 0;
    V3 axis = axis_of(o);
    V3 dir = to - from;
    double capture_height = o.thickness * 3.0;
    CylHit h = vol_clipped_cylinder(from, to, o.outer_radius, axis, capture_height);
    if (h.kind == CYL_ONE) hits[n++] = h.t1;
    if (h.kind == CYL_TWO) { hits[n++] = h.t1; hits[n++] = h.t2; }
    h = vol_clipped_cylinder(from, to, o.inner_radius, axis, capture_height);
    if (h.kind == CYL_ONE) hits[n++] = h.t1;
    if (h.kind == CYL_TWO) { hits[n++] = h.t1; hits[n++] = h.t2; }
    double caps[2] = {capture_height, -capture_height};
    for (int k = 0; k < 2; ++k) {
        double t;
        if (vol_cap(from, to, o.outer_radius, axis, caps[k], t)) {
            V3 p = from + t * dir;
            double r_sq = norm_sq(cross(p, axis));
            if (r_sq >= o.inner_radius * o.inner_radius) hits[n++] = t;
        }
    }
    std::stable_sort(hits, hits + n, total_less);
    if (n == 0) return CylHit{CYL_NONE, 0, 0};
    if (n == 1) return CylHit{CYL_ONE, hits[0], 0};
    return CylHit{CYL_TWO, hits[0], hits[1]};
}
void vol_get_uv(const GrrtObject& o, V3 p, double& u, double& v) {  // :140-152
    V3 e1 = v3(o.e1[0], o.e1[1], o.e1[2]), e2 = v3(o.e2[0], o.e2[1], o.e2[2]);
    double x = dot(p, e1), y = dot(p, e2);
    double rr = std::sqrt(x * x + y * y);
    double phi = std::atan2(y, x);
    double r = (rr - o.inner_radius) / (o.outer_radius - o.inner_radius);
    u = 0.5 + 0.5 * r * std::cos(phi);
    v = 0.5 + 0.5 * r * std::sin(phi);
}
bool vol_intersects(const GrrtObject& o, V3 ys, V3 ye, Intersection& out) {  // :506-578
    V3 direction = ye - ys;
    CylHit h = vol_cylinder(o, ys, ye);
    double t;
    if (h.kind == CYL_NONE || h.kind == CYL_PARALLEL) return false;
    if (h.kind == CYL_ONE) {
        if (h.t1 > MIN_INTERSECTION_T) t = h.t1;
        else return false;
    } else {
        if (h.t1 > MIN_INTERSECTION_T) t = h.t1;
        else if (h.t2 > MIN_INTERSECTION_T) t = h.t2;
        else return false;
    }
    V3 ip = ys + t * direction;
    if (!in01(t)) return false;
    vol_get_uv(o, ip, out.u, out.v);
    out.point = v4(0.0, ip.x, ip.y, ip.z);
    out.t = t;
    out.direction = direction;
    return true;
}
double vol_noise(const GrrtObject& o, V3 p) { return perlin_3d(o.perlin_perm, p.x, p.y, p.z); }  // :344-346
double vol_fbm(const GrrtObject& o, V3 x, double h) {  // :330-342
    double g = std::exp2(-h);
    double frequency = 4.0, amplitude = 1.0, t = 0.0;
    for (uint64_t i = 0; i < o.num_octaves; ++i) {
        t += amplitude * vol_noise(o, x * frequency);
        frequency *= 2.0;
        amplitude *= g;
    }
    return t;
}
double vol_density(const GrrtObject& o, V3 p) {  // :97-138
    V3 axis = axis_of(o);
    double h = std::fabs(dot(p, axis));
    double r = norm(cross(p, axis));
    if (r <= o.inner_radius || r >= o.outer_radius) return 0.0;
    double vertical_falloff = std::exp(-powi2(h / o.thickness));
    if (vertical_falloff < 0.001) return 0.0;
    double radial_base = std::pow(o.inner_radius / r, 1.5);
    double boundary_falloff = 1.0;
    boundary_falloff *= std::exp(-1.0 / rmax(powi2(o.outer_radius - r), 0.0001));
    boundary_falloff *= std::exp(-1.0 / rmax(powi2(r - o.inner_radius), 0.0001));
    V3 e1 = v3(o.e1[0], o.e1[1], o.e1[2]), e2 = v3(o.e2[0], o.e2[1], o.e2[2]);
    double x_local = dot(p, e1), y_local = dot(p, e2);
    double phi = std::atan2(y_local, x_local);
    double noise_phi_x = std::cos(phi) * o.noise_scale[1];
    double noise_phi_y = std::sin(phi) * o.noise_scale[1];
    V3 noise_p = v3(r * o.noise_scale[0], noise_phi_x, noise_phi_y);
    double n = vol_fbm(o, noise_p, 0.5);
    n += vol_noise(o, v3(r * 0.5, h * o.noise_scale[2], std::cos(phi))) * 0.5;
    double nn = rmax(n + o.noise_offset, 0.0) * o.density_multiplier;
    return nn * radial_base * vertical_falloff * boundary_falloff;
}
bool vol_does_exit(const GrrtObject& o, V3 p, V3 rd, double t) {  // :154-170
    V3 to = v3(p.x + rd.x * t, p.y + rd.y * t, p.z + rd.z * t);
    Intersection isec;
    if (vol_intersects(o, p, to, isec)) return isec.t > MIN_INTERSECTION_T;
    return false;
}
bool vol_precompute_exit(const GrrtObject& o, V3 ro, V3 rd, double& out) {  // :172-196
    double max_distance = o.vol_step_size * (double)o.vol_max_steps;
    V3 to = ro + rd * max_distance;
    CylHit h = vol_cylinder(o, ro, to);
    switch (h.kind) {
        case CYL_ONE:
            if (h.t1 > MIN_INTERSECTION_T) { out = h.t1 * max_distance; return true; }
            return false;
        case CYL_TWO:
            if (h.t1 > MIN_INTERSECTION_T) { out = h.t1 * max_distance; return true; }
            if (h.t2 > MIN_INTERSECTION_T) { out = h.t2 * max_distance; return true; }
            return false;
        default: return false;
    }
}
// :209-328; returns GRRT_RAYERR_* from compute_temperature (the caller maps it to (0,0,0,0))
int vol_raymarch(const SceneView& sv, const GrrtObject& o, V3 ro, V3 rd, const Frequency& fr, Cie& out,
                 uint64_t& samples, bool use_cached_exit = true) {
    double sigma_a = o.absorption, sigma_s = o.scattering;
    Cie accum = cie(0.0, 0.0, 0.0, 0.0);
    double transparency = 1.0, alpha_weighted_sum = 0.0, alpha_weight_total = 0.0;
    double d_s = o.vol_step_size, d_o = 0.0;
    double exit_distance = 0.0;
    // raymarch_constant_step_internal(.., use_cached_exit): the production path passes true (:198-207)
    bool cached = use_cached_exit && vol_precompute_exit(o, ro, rd, exit_distance);
    V3 axis = axis_of(o);
    for (uint64_t i = 0; i < o.vol_max_steps; ++i) {
        V3 p = ro + rd * d_o;
        d_o += d_s;
        ++samples;
        double density = vol_density(o, p);
        if (density > 0.0) {
            double sample_attenuation = std::exp(-d_s * density * (sigma_a + sigma_s));
            transparency *= sample_attenuation;
            double u_t, u_phi;
            if (orbit_coefficients(sv.g, CART, v4(0.0, p.x, p.y, p.z), u_t, u_phi)) {
                double emitter_energy = u_t * fr.p_t + u_phi * fr.p_phi;
                double redshift = fr.observer_energy / emitter_energy;
                double r_dist = norm(cross(p, axis));
                double temperature;
                int err = compute_temperature(o.temperature, r_dist, temperature);
                if (err) return err;
                double u, v;
                vol_get_uv(o, p, u, v);
                Cie light = texture_color(*sv.s, o.texture, u, v, temperature, redshift);
                double travel_density = d_s;
                double light_attenuation = std::exp(-density * travel_density * (sigma_a + sigma_s));
                double intensity_factor = powi4(temperature / o.brightness_reference_temperature);
                double emission_weight = transparency * light_attenuation * sigma_s * density * d_s;
                double f = emission_weight * intensity_factor;
                double alpha_sample_weight = density * d_s;
                alpha_weighted_sum += rclamp(light.alpha, 0.0, 1.0) * alpha_sample_weight;
                alpha_weight_total += alpha_sample_weight;
                accum.x += light.x * f;
                accum.y += light.y * f;
                accum.z += light.z * f;
            }
        }
        bool exited = cached ? (d_o >= exit_distance) : vol_does_exit(o, p, rd, d_s);
        if (exited) break;
    }
    double physical_opacity = 1.0 - transparency;
    double texture_alpha = alpha_weight_total > 0.0 ? alpha_weighted_sum / alpha_weight_total : 1.0;
    accum.alpha = physical_opacity * texture_alpha;
    out = accum;
    return 0;
}

// objects.rs:27-44
StepRec step_at_intersection(const Geom& g, const StepRec& ys, const StepRec& ye, V4 point, double t) {
    double s = 1.0 - t;
    StepRec r;
    r.t = s * ys.t + t * ye.t;
    r.step = ys.step;
    r.x = cartesian_to_chart(g, g.chart(), point);
    r.p = v4(s * ys.p[0] + t * ye.p[0], s * ys.p[1] + t * ye.p[1], s * ys.p[2] + t * ye.p[2],
             s * ys.p[3] + t * ye.p[3]);
    return r;
}

// Objects::intersects, objects.rs:65-120.  Returns GRRT_RAYERR_* (0 ok); `found` tells whether
// a colour was produced.
int objects_intersect(const SceneView& sv, const StepRec& y_start, const StepRec& y_end, const Frequency& fr,
                      bool& found, Cie& color, uint64_t& samples) {
    const Geom& g = sv.g;
    found = false;
    double shortest = std::numeric_limits<double>::max();
    Chart c = g.chart();
    V3 ys = spatial_cartesian(g, c, y_start.x);
    V3 ye = spatial_cartesian(g, c, y_end.x);
    for (uint32_t k = 0; k < sv.s->n_objects; ++k) {
        const GrrtObject& o = sv.s->objects[k];
        Intersection isec;
        bool hit = false;
        switch (o.kind) {
            case GRRT_OBJ_SPHERE: hit = sphere_intersects(o, ys, ye, isec); break;
            case GRRT_OBJ_DISC: hit = disc_intersects(o, ys, ye, isec); break;
            default: hit = vol_intersects(o, ys, ye, isec); break;
        }
        if (!hit) continue;
        V3 ip = v3(isec.point[1], isec.point[2], isec.point[3]);
        double distance = norm(ip - ys);
        if (distance < shortest) {
            shortest = distance;
            StepRec st = step_at_intersection(g, y_start, y_end, isec.point, isec.t);
            // energy_of_emitter: sphere.rs:141-150 (static), disc.rs:101-110 / volumetric_disc.rs:603-612 (orbit)
            V4 velocity;
            if (o.kind == GRRT_OBJ_SPHERE) {
                velocity = stationary_velocity(g, st.x);
            } else {
                if (!circular_orbit_velocity(g, st.x, velocity)) return GRRT_RAYERR_NO_CIRCULAR_ORBIT;
            }
            double emitter_energy = inner_product(g, st.x, velocity, st.p);
            double redshift = (g.sig0() * fr.observer_energy) / (g.sig0() * emitter_energy);  // redshift.rs:36-38,69-73
            double temperature;
            if (o.kind == GRRT_OBJ_SPHERE) {
                temperature = o.sphere_temperature;
            } else {
                double r = radial_coordinate(g, CART, isec.point);
                int err = compute_temperature(o.temperature, r, temperature);
                if (err) return err;
            }
            if (o.kind == GRRT_OBJ_VOLUMETRIC_DISC) {  // volumetric_disc.rs:580-601
                V3 rd = normalize(isec.direction);
                Cie c2;
                int err = vol_raymarch(sv, o, ip, rd, fr, c2, samples);
                color = err ? cie(0.0, 0.0, 0.0, 0.0) : c2;
            } else {
                color = texture_color(*sv.s, o.texture, isec.u, isec.v, temperature, redshift);
            }
            found = true;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// per-ray pipeline: Integrator::integrate (integrator.rs:78-174, 203-268) fused with
// Scene::color_of_ray (scene.rs:114-231).  The reference stores every step and then walks
// steps.windows(2); the pairs are processed here as they are produced -- same values, and any
// Err(..) of either phase makes the whole ray an error, exactly as `?` does there.
// ---------------------------------------------------------------------------------------
struct RayResult {
    Cie color;
    int ray_class;
    int stop;    // GRRT_STOP_*
    int error;   // GRRT_RAYERR_*
    uint64_t steps, attempts, samples;
    uint32_t n_hits;
};

int should_stop(const Geom& g, const Y8& y, uint64_t step_index, uint64_t max_steps, double max_radius_sq) {
    for (int i = 0; i < 4; ++i)
        if (!std::isfinite(y[i])) return GRRT_STOP_NAN;
    if (inside_horizon(g, y)) return GRRT_STOP_HORIZON;
    if (closed_orbit(g, y, step_index, max_steps)) return GRRT_STOP_CLOSED_ORBIT;
    V4 c = to_cartesian(g, g.chart(), v4(y[0], y[1], y[2], y[3]));  // scene.rs:48-69
    if (c[1] * c[1] + c[2] * c[2] + c[3] * c[3] > max_radius_sq) return GRRT_STOP_CELESTIAL;
    for (int i = 4; i < 8; ++i)
        if (!std::isfinite(y[i])) return GRRT_STOP_NAN;
    return GRRT_STOP_NONE;
}

template <class StepSink>
RayResult trace_ray(const SceneView& sv, const Ray& ray, bool shade, StepSink&& sink) {
    const Geom& g = sv.g;
    const GrrtScene& sc = *sv.s;
    RayResult res;
    std::memset(&res, 0, sizeof res);
    res.color = cie(0.0, 0.0, 0.0, 1.0);
    res.ray_class = GRRT_CLASS_ESCAPED;

    Solver solver = make_solver(g, ray);
    auto f = [&](const Y8& yy) { return geodesic(solver, yy); };
    Y8 y = create_initial_state(solver, ray);
    double t = 0.0;
    StepRec prev;
    prev.x = v4(y[0], y[1], y[2], y[3]);
    prev.p = momentum_from_state(solver, y);
    prev.t = t;
    prev.step = 0;
    sink(prev);

    // redshift.rs:45-60
    Frequency fr;
    V4 cam_vel = v4(sc.camera.velocity[0], sc.camera.velocity[1], sc.camera.velocity[2], sc.camera.velocity[3]);
    fr.observer_energy = inner_product(g, ray.pos, cam_vel, ray.mom);
    fr.p_t = inner_product(g, ray.pos, v4(1.0, 0.0, 0.0, 0.0), ray.mom);
    fr.p_phi = inner_product(g, ray.pos, axial_killing_vector(g, ray.pos), ray.mom);

    std::vector<Cie> intersections;
    double object_opacity = 0.0;
    double max_radius_sq = sc.integration.max_radius * sc.integration.max_radius;
    uint64_t max_steps = sc.integration.max_steps;
    double h = sc.integration.step_size;
    int stop = GRRT_STOP_NONE;
    for (uint64_t i = 1; i < max_steps; ++i) {
        Y8 y_new;
        double h_taken, h_next;
        if (!rkf45(y, h, sc.integration.epsilon, f, y_new, h_taken, h_next, res.attempts)) {
            res.error = GRRT_RAYERR_RKF45_RETRIES;
            res.stop = GRRT_STOP_ERROR;
            return res;
        }
        y = y_new;
        t += h_taken;
        h = h_next;
        StepRec cur;
        cur.x = v4(y[0], y[1], y[2], y[3]);
        cur.p = momentum_from_state(solver, y);
        cur.t = t;
        cur.step = i;
        sink(cur);
        res.steps = i;
        if (shade) {
            bool found;
            Cie c;
            int err = objects_intersect(sv, prev, cur, fr, found, c, res.samples);
            if (err) {
                res.error = err;
                res.stop = GRRT_STOP_ERROR;
                return res;
            }
            if (found) {
                intersections.push_back(c);
                double alpha = rclamp(c.alpha, 0.0, 1.0);
                object_opacity = alpha + object_opacity * (1.0 - alpha);
            }
        }
        prev = cur;
        stop = should_stop(g, y, i, max_steps, max_radius_sq);
        if (stop != GRRT_STOP_NONE) break;
    }
    res.stop = stop;
    res.n_hits = (uint32_t)intersections.size();
    if (!shade) return res;

    int ray_class;
    switch (stop) {  // scene.rs:157-205
        case GRRT_STOP_HORIZON:
        case GRRT_STOP_CLOSED_ORBIT:
            intersections.push_back(cie(0.0, 0.0, 0.0, 1.0));
            ray_class = GRRT_CLASS_CAPTURED;
            break;
        case GRRT_STOP_CELESTIAL: {
            V3 sph = get_as_spherical(g.chart(), prev.x);  // scene.rs:222-231
            double theta = sph.y, phi = sph.z;
            double u = (PI + phi) / (2.0 * PI);
            double v = theta / PI;
            // redshift.rs:31-34, 62-67: static emitter at the last step
            V4 vel = stationary_velocity(g, prev.x);
            double emitter_energy = inner_product(g, prev.x, vel, prev.p);
            double redshift = (g.sig0() * fr.observer_energy) / (g.sig0() * emitter_energy);
            intersections.push_back(texture_color(sc, sc.celestial, 1.0 - u, v, sc.celestial_temperature, redshift));
            ray_class = GRRT_CLASS_ESCAPED;
            break;
        }
        default: ray_class = GRRT_CLASS_CAPTURED; break;  // NaN or budget exhausted
    }
    Cie result = cie(0.0, 0.0, 0.0, 1.0);
    for (size_t k = intersections.size(); k-- > 0;) result = blend(result, intersections[k]);
    if (object_opacity >= sc.adaptive.object_hit_opacity_threshold) ray_class = GRRT_CLASS_HIT;
    res.color = result;
    res.ray_class = ray_class;
    return res;
}

SceneView view_of(const GrrtScene* s) {
    SceneView sv;
    sv.s = s;
    sv.g.kind = s->geometry.kind;
    sv.g.radius = s->geometry.radius;
    sv.g.a = s->geometry.a;
    sv.g.heps = s->geometry.horizon_epsilon;
    sv.g.kerr_mode = s->kerr_rhs_mode;
    return sv;
}

// ---------------------------------------------------------------------------------------
// adaptive sampling (rendering/raytracer.rs:59-159, 386-458)
// ---------------------------------------------------------------------------------------
uint64_t mix64(uint64_t z) {  // raytracer.rs:132-136
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
double hash_pixel_samples(int64_t row, int64_t col, uint64_t k) {  // :138-143
    uint64_t z = mix64((uint64_t)row + mix64((uint64_t)col + mix64(k)));
    return (double)(z >> 11) * (1.0 / (double)(1ull << 53));
}
void stratified_sample_offset(int64_t row, int64_t col, uint64_t sr, uint64_t sc, uint64_t n, double& dx, double& dy) {
    uint64_t sample_index = sr * n + sc;  // :145-159
    dx = ((double)sc + hash_pixel_samples(row, col, 2 * sample_index)) / (double)n;
    dy = ((double)sr + hash_pixel_samples(row, col, 2 * sample_index + 1)) / (double)n;
}
bool should_supersample_pair(const double* p, int pc, const double* q, int qc, const GrrtAdaptive& cfg,
                             double minimum_luminance) {  // :87-105
    if (pc != qc) return true;
    if (cfg.exclude_background_contrast && pc == GRRT_CLASS_ESCAPED) return false;
    bool visible = rmax(p[1], q[1]) > minimum_luminance;
    double lc = std::fabs(p[1] - q[1]) / (p[1] + q[1] + 1e-4);
    double oc = std::fabs(p[3] - q[3]);
    return visible && (lc > cfg.luminance_contrast_threshold || oc > cfg.opacity_contrast_threshold);
}
double resolve_minimum_luminance(const GrrtAdaptive& cfg, const double* xyza, size_t n) {  // :118-129
    if (cfg.has_minimum_luminance) return cfg.minimum_luminance;
    if (n == 0) return 0.0;
    std::vector<double> lum(n);
    for (size_t i = 0; i < n; ++i) lum[i] = xyza[4 * i + 1];
    size_t index = (size_t)sat_u64((double)(n - 1) * 0.99);
    std::nth_element(lum.begin(), lum.begin() + index, lum.end(), total_less);
    return 1e-3 * lum[index];
}

}  // namespace

// =========================================================================================
// C entry points (ctypes from tests/, bench.py --impl reference, __graft_entry__.smoke)
// =========================================================================================
extern "C" {

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// ---- scene assembly helpers (cli/*.rs, camera.rs, texture.rs, temperature.rs) ----------
// camera position: Cartesian (main.rs:92-97) -> native chart (cli/schwarzschild.rs:24, kerr.rs:23, kerr_bl.rs:23)
void orc_camera_position_native(const GrrtGeometry* geom, const double cart[3], double out[4]) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    V4 c = v4(0.0, cart[0], cart[1], cart[2]);
    V4 r = cartesian_to_chart(g, g.chart(), c);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
}
// resolve_camera_velocity + assert_future_directed, cli/shared.rs:23-86
int orc_resolve_camera_velocity(const GrrtGeometry* geom, const double pos[4], int mode, const double explicit_u[4],
                                double out[4]) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    V4 p = v4(pos[0], pos[1], pos[2], pos[3]);
    V4 u;
    if (mode == GRRT_CAMVEL_STATIC) u = stationary_velocity(g, p);
    else if (mode == GRRT_CAMVEL_ZAMO) u = zamo_velocity(g, p);
    else {
        u = v4(explicit_u[0], explicit_u[1], explicit_u[2], explicit_u[3]);
        double n = inner_product(g, p, u, u);
        if (std::fabs(n - g.sig0()) > 1e-6) return GRRT_ERR_INVALID_CONFIGURATION;
    }
    V4 tt = tetrad_at(g, p).t;
    double orientation = g.sig0() * inner_product(g, p, tt, u);
    if (!(std::isfinite(u[0]) && 1.0 * orientation > 0.0)) return GRRT_ERR_INVALID_CONFIGURATION;
    for (int i = 0; i < 4; ++i) out[i] = u[i];
    return 0;
}
int orc_camera_new(const GrrtGeometry* geom, const double pos[4], const double vel[4], double alpha, int64_t rows,
                   int64_t cols, double phi, double theta, double psi, GrrtCamera* out) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    return camera_new(g, v4(pos[0], pos[1], pos[2], pos[3]), v4(vel[0], vel[1], vel[2], vel[3]), alpha, rows, cols,
                      phi, theta, psi, out);
}
void orc_tetrad_at(const GrrtGeometry* geom, const double pos[4], double out[16]) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    Tetrad T = tetrad_at(g, v4(pos[0], pos[1], pos[2], pos[3]));
    for (int i = 0; i < 4; ++i) {
        out[i] = T.t[i];
        out[4 + i] = T.x[i];
        out[8 + i] = T.y[i];
        out[12 + i] = T.z[i];
    }
}
double orc_inner_product(const GrrtGeometry* geom, const double pos[4], const double v[4], const double w[4]) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    return inner_product(g, v4(pos[0], pos[1], pos[2], pos[3]), v4(v[0], v[1], v[2], v[3]), v4(w[0], w[1], w[2], w[3]));
}
void orc_camera_ray(const GrrtCamera* cam, int64_t row, int64_t col, int use_offset, double dx, double dy,
                    double pos[4], double mom[4]) {
    Ray r = use_offset ? camera_ray_offset(*cam, row, col, dx, dy) : camera_ray(*cam, row, col);
    for (int i = 0; i < 4; ++i) {
        pos[i] = r.pos[i];
        mom[i] = r.mom[i];
    }
}
// BlackBodyMapper::new, texture.rs:121-138
void orc_build_blackbody_lut(double* log_t, double* xyz) {
    const int N = GRRT_BLACKBODY_LUT_N;
    double min_log = std::log10(10.0), max_log = std::log10(10000000.0);
    double step = (max_log - min_log) / (double)(N - 1);
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < N; ++i) {
        double lt = min_log + (double)i * step;
        double t = std::pow(10.0, lt);
        Cie c = integrate_blackbody_xyz(t, 1.0);
        log_t[i] = lt;
        xyz[3 * i] = c.x;
        xyz[3 * i + 1] = c.y;
        xyz[3 * i + 2] = c.z;
    }
}
void orc_blackbody_xyz(double temperature, double redshift, double out[4]) {
    Cie c = integrate_blackbody_xyz(temperature, redshift);
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.alpha;
}
double orc_r_isco(double r_s, double a) { return co_r_isco(r_s, a); }
int orc_circular_orbit(double r_s, double a, double r, double out[2]) {
    return co_killing(r_s, a, r, out[0], out[1]) ? 0 : GRRT_RAYERR_NO_CIRCULAR_ORBIT;
}

}  // extern "C"

namespace {
// KerrTemperatureComputer, temperature.rs:45-195
struct KerrTemp {
    double a, radius, r_isco, m_dot;
    bool l(double r, double& o) const { return co_conserved_angular_momentum(radius, a, r, o); }
    bool e(double r, double& o) const { return co_conserved_energy(radius, a, r, o); }
    double omega(double r) const { return co_angular_velocity(radius, a, r); }
    bool d_l_dr(double r, double& o) const {  // :127-139
        double h = 1e-6 * rmax(r, 1.0);
        double lp, l0;
        if (r - h < r_isco) {
            if (!l(r + h, lp) || !l(r, l0)) return false;
            o = (lp - l0) / h;
        } else {
            if (!l(r + h, lp) || !l(r - h, l0)) return false;
            o = (lp - l0) / (2.0 * h);
        }
        return true;
    }
    double d_omega_dr(double r) const {  // :141-145
        double h = 1e-10;
        return (omega(r + h) - omega(r - h)) / (2.0 * h);
    }
    int integral(double r, double& o) const {  // :177-190
        const int N = 1000;
        double dr = (r - r_isco) / (double)N;
        double acc = 0.0;
        for (int i = 0; i < N; ++i) {
            double rp = r_isco + ((double)i + 0.5) * dr;
            double ep, lp, dl;
            if (!e(rp, ep) || !l(rp, lp)) return GRRT_ERR_LUT;
            double om = omega(rp);
            if (!d_l_dr(rp, dl)) return GRRT_ERR_LUT;
            acc += (ep - om * lp) * dl * dr;
        }
        o = acc;
        return 0;
    }
    int prefactor(double r, double& o) const {  // :159-175
        double ee, ll;
        if (!e(r, ee) || !l(r, ll)) return GRRT_ERR_LUT;
        double om = omega(r);
        double denominator = (r * r) * powi2(ee - om * ll);
        if (std::fabs(denominator) < 1e-20) return GRRT_ERR_LUT;
        o = d_omega_dr(r) / denominator;
        return 0;
    }
    int f(double r, double& o) const {  // :147-157
        double in, pre;
        int err = integral(r, in);
        if (err) return err;
        err = prefactor(r, pre);
        if (err) return err;
        double coefficient = -m_dot / (PI * radius * radius);
        o = coefficient * pre * in;
        return 0;
    }
};
}  // namespace

extern "C" {

// KerrTemperatureComputer::new, temperature.rs:45-118
int orc_build_temperature_lut(double temperature, double outer_radius, double a, double radius, double* lut_r,
                              double* lut_t, double* r_isco_out) {
    const int N = GRRT_TEMPERATURE_LUT_N;
    double a_abs = std::fabs(a);
    double r_isco = co_r_isco(radius, a_abs);
    double eff_outer = outer_radius <= r_isco ? r_isco + rmax(1e-6, std::fabs(r_isco) * 1e-9) : outer_radius;
    KerrTemp tc{a_abs, radius, r_isco, 1.0};
    double max_f = 0.0, max_r = 0.0;
    double dr = (eff_outer - r_isco) / 10.0;
    for (int i = 0; i < 10; ++i) {
        double r = r_isco + ((double)i + 0.5) * dr;
        double fv;
        int err = tc.f(r, fv);
        if (err) return err;
        if (max_f < fv) {
            max_f = fv;
            max_r = r;
        }
    }
    double in, pre;
    int err = tc.integral(max_r, in);
    if (err) return err;
    err = tc.prefactor(max_r, pre);
    if (err) return err;
    double coefficient = -1.0 / (PI * radius * radius);
    double f = 1.0 * std::pow(temperature, 4.0);
    tc.m_dot = f / (coefficient * pre * in);
    double step = (eff_outer - r_isco) / (double)(N - 1);
    int fail = 0;
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < N; ++i) {
        double r = r_isco + (double)i * step;
        double fv = 0.0;
        int e2 = tc.f(r, fv);
        if (e2) {
#pragma omp atomic write
            fail = e2;
        }
        lut_r[i] = r;
        lut_t[i] = std::pow(rmax(fv / 1.0, 0.0), 0.25);
    }
    *r_isco_out = r_isco;
    return fail;
}
int orc_compute_temperature(const GrrtTemperature* tc, double radius, double* out) {
    return compute_temperature(*tc, radius, *out);
}
void orc_perlin_permutation(uint32_t seed, uint8_t* out) { perlin_permutation(seed, out); }
double orc_perlin(const uint8_t* perm, double x, double y, double z) { return perlin_3d(perm, x, y, z); }
// VolumetricDisc::new axis / e1 / e2, volumetric_disc.rs:61-73
void orc_volumetric_basis(const double axis_in[3], double axis[3], double e1[3], double e2[3]) {
    V3 ax = v3(axis_in[0], axis_in[1], axis_in[2]);
    if (norm_sq(ax) <= 2.220446049250313e-16) ax = v3(0.0, 0.0, 1.0);
    else ax = normalize(ax);
    V3 b = std::fabs(ax.x) > 0.9 ? v3(0.0, 1.0, 0.0) : v3(1.0, 0.0, 0.0);
    V3 E1 = normalize(cross(b, ax));
    V3 E2 = normalize(cross(ax, E1));
    axis[0] = ax.x; axis[1] = ax.y; axis[2] = ax.z;
    e1[0] = E1.x; e1[1] = E1.y; e1[2] = E1.z;
    e2[0] = E2.x; e2[1] = E2.y; e2[2] = E2.z;
}
double orc_volumetric_density(const GrrtObject* o, const double p[3]) { return vol_density(*o, v3(p[0], p[1], p[2])); }
// ---- extra probes for the reference's unit tests (tests/test_oracle_kat.py) ----
void orc_ks_metrics(double radius, double a, double x, double y, double z, double cov[16], double contra[16]) {
    M4 c = ks_metric(radius, a, x, y, z), k = ks_metric_contravariant(radius, a, x, y, z);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            cov[4 * i + j] = c.m[i][j];
            contra[4 * i + j] = k.m[i][j];
        }
}
void orc_axial_killing_vector(const GrrtGeometry* geom, const double pos[4], double out[4]) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    V4 v = axial_killing_vector(g, v4(pos[0], pos[1], pos[2], pos[3]));
    for (int i = 0; i < 4; ++i) out[i] = v[i];
}
int orc_circular_orbit_velocity(const GrrtGeometry* geom, const double pos[4], double out[4]) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    V4 v;
    if (!circular_orbit_velocity(g, v4(pos[0], pos[1], pos[2], pos[3]), v)) return 1;
    for (int i = 0; i < 4; ++i) out[i] = v[i];
    return 0;
}
// Sphere::intersects / VolumetricDisc::intersects on a Cartesian chord: 1 = hit (world-space point, chord parameter)
int orc_object_intersects(const GrrtObject* o, const double a[3], const double b[3], double hit[3], double* t) {
    Intersection isec;
    bool ok = false;
    V3 ys = v3(a[0], a[1], a[2]), ye = v3(b[0], b[1], b[2]);
    if (o->kind == GRRT_OBJ_SPHERE) ok = sphere_intersects(*o, ys, ye, isec);
    else if (o->kind == GRRT_OBJ_VOLUMETRIC_DISC) ok = vol_intersects(*o, ys, ye, isec);
    else return -1;
    if (ok) {
        hit[0] = isec.point[1]; hit[1] = isec.point[2]; hit[2] = isec.point[3];
        if (t) *t = isec.t;
    }
    return ok ? 1 : 0;
}
// VolumetricDisc::raymarch_constant_step_internal (volumetric_disc.rs:209-328) with either exit test
int orc_volumetric_raymarch(const GrrtScene* s, uint32_t object_index, const double ro[3], const double rd[3],
                            const double freq[3], int use_cached_exit, double xyza[4], uint64_t* samples) {
    if (object_index >= s->n_objects || s->objects[object_index].kind != GRRT_OBJ_VOLUMETRIC_DISC) return -1;
    SceneView sv = view_of(s);
    Frequency fr{freq[0], freq[1], freq[2]};
    Cie c;
    uint64_t n = 0;
    int rc = vol_raymarch(sv, s->objects[object_index], v3(ro[0], ro[1], ro[2]), v3(rd[0], rd[1], rd[2]), fr, c, n,
                          use_cached_exit != 0);
    xyza[0] = c.x; xyza[1] = c.y; xyza[2] = c.z; xyza[3] = c.alpha;
    if (samples) *samples = n;
    return rc;
}
void orc_srgb_to_xyz(uint8_t r, uint8_t g, uint8_t b, double out[4]) {
    Cie c = srgb_to_xyz(r, g, b);
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.alpha;
}
void orc_blend(const double self[4], const double other[4], double out[4]) {
    Cie r = blend(cie(self[0], self[1], self[2], self[3]), cie(other[0], other[1], other[2], other[3]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.alpha;
}
void orc_texture_color(const GrrtScene* s, const GrrtTexture* t, double u, double v, double temperature,
                       double redshift, double out[4]) {
    Cie c = texture_color(*s, *t, u, v, temperature, redshift);
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.alpha;
}
double orc_hash_pixel_samples(int64_t row, int64_t col, uint64_t k) { return hash_pixel_samples(row, col, k); }
void orc_stratified_sample_offset(int64_t row, int64_t col, uint64_t sr, uint64_t sc, uint64_t n, double* dx, double* dy) {
    stratified_sample_offset(row, col, sr, sc, n, *dx, *dy);
}
int orc_should_supersample_pair(const double p[4], int pc, const double q[4], int qc, const GrrtAdaptive* cfg,
                                double minimum_luminance) {
    return should_supersample_pair(p, pc, q, qc, *cfg, minimum_luminance) ? 1 : 0;
}

// generic RKF45 on y'' = 2 (runge_kutta.rs:190-239 known-answer test)
void orc_rkf45_simple_equation(double t_end, double eps, double step, double out[3]) {
    Y8 y = Y8{{1.0, 2.0, 0, 0, 0, 0, 0, 0}};
    auto f = [](const Y8& v) { return Y8{{v[1], 2.0, 0, 0, 0, 0, 0, 0}}; };
    double t = 0.0, h = step;
    uint64_t attempts = 0;
    while (t <= t_end) {
        Y8 yn;
        double ht, hn;
        if (!rkf45(y, h, eps, f, yn, ht, hn, attempts)) break;
        y = yn;
        t += ht;
        h = hn;
    }
    out[0] = t;
    out[1] = y[0];
    out[2] = y[1];
}
int orc_should_stop(const GrrtGeometry* geom, const double y[8], uint64_t step_index, uint64_t max_steps,
                    double max_radius) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    Y8 yy;
    for (int i = 0; i < 8; ++i) yy[i] = y[i];
    return should_stop(g, yy, step_index, max_steps, max_radius * max_radius);
}
void orc_geodesic_rhs(const GrrtScene* s, const double pos[4], const double mom[4], const double y[8], double out[8]) {
    SceneView sv = view_of(s);
    Ray ray{v4(pos[0], pos[1], pos[2], pos[3]), v4(mom[0], mom[1], mom[2], mom[3]), 0, 0};
    Solver solver = make_solver(sv.g, ray);
    Y8 yy;
    for (int i = 0; i < 8; ++i) yy[i] = y[i];
    Y8 d = geodesic(solver, yy);
    for (int i = 0; i < 8; ++i) out[i] = d[i];
}
void orc_solver_constants(const GrrtScene* s, const double pos[4], const double mom[4], double out[3], double y0[8]) {
    SceneView sv = view_of(s);
    Ray ray{v4(pos[0], pos[1], pos[2], pos[3]), v4(mom[0], mom[1], mom[2], mom[3]), 0, 0};
    Solver solver = make_solver(sv.g, ray);
    out[0] = solver.e; out[1] = solver.l_z; out[2] = solver.q;
    Y8 y = create_initial_state(solver, ray);
    for (int i = 0; i < 8; ++i) y0[i] = y[i];
}

// Integrator::integrate with every step stored (raytracer.rs:499-507 / cli/shared.rs:107-129)
int orc_trace_ray(const GrrtScene* s, const double pos[4], const double mom[4], GrrtStep* steps_out, uint64_t cap,
                  uint64_t* n_out, int32_t* stop_reason) {
    SceneView sv = view_of(s);
    Ray ray{v4(pos[0], pos[1], pos[2], pos[3]), v4(mom[0], mom[1], mom[2], mom[3]), 0, 0};
    uint64_t n = 0;
    RayResult r = trace_ray(sv, ray, false, [&](const StepRec& st) {
        if (n < cap && steps_out) {
            for (int i = 0; i < 4; ++i) {
                steps_out[n].x[i] = st.x[i];
                steps_out[n].p[i] = st.p[i];
            }
            steps_out[n].t = st.t;
            steps_out[n].step = st.step;
        }
        ++n;
    });
    *n_out = n;
    *stop_reason = r.stop;
    return r.error;
}
int orc_trace_pixel_ray(const GrrtScene* s, int64_t row, int64_t col, GrrtStep* steps_out, uint64_t cap,
                        uint64_t* n_out, int32_t* stop_reason) {
    Ray ray = camera_ray(s->camera, row, col);
    return orc_trace_ray(s, ray.pos.v, ray.mom.v, steps_out, cap, n_out, stop_reason);
}
// position (any chart) -> Cartesian, for the CSV of ray.rs:35-54
void orc_step_to_cartesian(const GrrtGeometry* geom, const double x[4], double out[4]) {
    Geom g{geom->kind, geom->radius, geom->a, geom->horizon_epsilon, 0};
    V4 c = to_cartesian(g, g.chart(), v4(x[0], x[1], x[2], x[3]));
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}

// Scene::color_of_ray for one camera ray (scene.rs:114-220)
int orc_color_of_ray(const GrrtScene* s, int64_t row, int64_t col, int use_offset, double dx, double dy, double xyza[4],
                     int32_t* ray_class, int32_t* stop_reason, uint64_t* steps, uint64_t* attempts) {
    SceneView sv = view_of(s);
    Ray ray = use_offset ? camera_ray_offset(s->camera, row, col, dx, dy) : camera_ray(s->camera, row, col);
    RayResult r = trace_ray(sv, ray, true, [](const StepRec&) {});
    xyza[0] = r.color.x; xyza[1] = r.color.y; xyza[2] = r.color.z; xyza[3] = r.color.alpha;
    *ray_class = r.ray_class;
    *stop_reason = r.stop;
    if (steps) *steps = r.steps;
    if (attempts) *attempts = r.attempts;
    return r.error;
}

// Raytracer::render_section_to_cie_buffer (raytracer.rs:177-458).  `pass_mask`: bit0 = run the
// adaptive pass (otherwise stop after the raw 1-spp pass, as render_section_to_cie_buffer_raw).
int orc_render(const GrrtScene* s, uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col,
               double* xyza_out, uint8_t* class_out, uint8_t* stop_out, uint32_t* steps_out, GrrtStats* stats,
               int base_only) {
    SceneView sv = view_of(s);
    if (to_row < from_row || to_col < from_col) return GRRT_ERR_INVALID_ARGUMENT;
    const uint32_t W = to_col - from_col, H = to_row - from_row;
    const size_t n = (size_t)W * H;
    std::vector<uint8_t> cls(n, GRRT_CLASS_ESCAPED);
    GrrtStats st;
    std::memset(&st, 0, sizeof st);
    uint64_t acc_steps = 0, acc_att = 0, acc_samples = 0, acc_err = 0;
    uint64_t hist[6] = {0, 0, 0, 0, 0, 0};
    // raytracer.rs:195-244
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : acc_steps, acc_att, acc_samples, acc_err, hist[:6])
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint32_t y = (uint32_t)(i / W), x = (uint32_t)(i % W);
        Ray ray = camera_ray(s->camera, (int64_t)(y + from_row), (int64_t)(x + from_col));
        RayResult r = trace_ray(sv, ray, true, [](const StepRec&) {});
        double* o = xyza_out + 4 * i;
        if (r.error) {  // pixel keeps its initial value (raytracer.rs:204-210, 230-239)
            o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 1.0;
            cls[i] = GRRT_CLASS_ESCAPED;
            ++acc_err;
        } else {
            o[0] = r.color.x; o[1] = r.color.y; o[2] = r.color.z; o[3] = r.color.alpha;
            cls[i] = (uint8_t)r.ray_class;
        }
        if (stop_out) stop_out[i] = (uint8_t)r.stop;
        if (steps_out) steps_out[i] = (uint32_t)r.steps;
        acc_steps += r.steps;
        acc_att += r.attempts;
        acc_samples += r.samples;
        hist[r.stop] += 1;
    }
    st.rays_base = n;
    const GrrtAdaptive& ad = s->adaptive;
    if (!base_only && (ad.enabled || ad.show_sampling_mask)) {  // raytracer.rs:184, 257-318
        double min_lum = resolve_minimum_luminance(ad, xyza_out, n);
        st.minimum_luminance = min_lum;
        std::vector<uint32_t> sel;
        static const int SH[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
        for (uint32_t row = 0; row < H; ++row)
            for (uint32_t col = 0; col < W; ++col) {
                size_t pi = (size_t)row * W + col;
                for (int k = 0; k < 8; ++k) {
                    int nr = (int)row + SH[k][0], nc = (int)col + SH[k][1];
                    if (nr < 0 || nr >= (int)H || nc < 0 || nc >= (int)W) continue;
                    size_t ni = (size_t)nr * W + nc;
                    if (should_supersample_pair(xyza_out + 4 * pi, cls[pi], xyza_out + 4 * ni, cls[ni], ad, min_lum)) {
                        sel.push_back((uint32_t)pi);
                        break;
                    }
                }
            }
        st.pixels_supersampled = sel.size();
        if (ad.show_sampling_mask) {
            for (uint32_t pi : sel)
                for (int c = 0; c < 4; ++c) xyza_out[4 * (size_t)pi + c] = ad.mask_xyza[c];
        } else {  // supersample, raytracer.rs:320-384
            const uint64_t N = ad.samples_per_axis;
            uint64_t ss_rays = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : acc_steps, acc_att, acc_samples, acc_err, ss_rays, hist[:6])
            for (int64_t j = 0; j < (int64_t)sel.size(); ++j) {
                uint32_t pi = sel[j];
                int64_t row = (int64_t)(pi / W) + from_row, col = (int64_t)(pi % W) + from_col;
                Cie acc = cie(0.0, 0.0, 0.0, 0.0);
                uint32_t valid = 0;
                for (uint64_t sr = 0; sr < N; ++sr)
                    for (uint64_t sc2 = 0; sc2 < N; ++sc2) {
                        double dx, dy;
                        stratified_sample_offset(row, col, sr, sc2, N, dx, dy);
                        Ray ray = camera_ray_offset(s->camera, row, col, dx, dy);
                        RayResult r = trace_ray(sv, ray, true, [](const StepRec&) {});
                        ++ss_rays;
                        acc_steps += r.steps;
                        acc_att += r.attempts;
                        acc_samples += r.samples;
                        hist[r.stop] += 1;
                        if (r.error) {
                            ++acc_err;
                            continue;
                        }
                        acc = cie(acc.x + r.color.x, acc.y + r.color.y, acc.z + r.color.z, acc.alpha + r.color.alpha);
                        ++valid;
                    }
                if (valid > 0) {
                    double inv = 1.0 / (double)valid;
                    double* o = xyza_out + 4 * (size_t)pi;
                    o[0] = acc.x * inv; o[1] = acc.y * inv; o[2] = acc.z * inv; o[3] = acc.alpha * inv;
                }
            }
            st.rays_supersample = ss_rays;
        }
    }
    st.accepted_steps = acc_steps;
    st.attempts = acc_att;
    st.raymarch_samples = acc_samples;
    st.ray_errors = acc_err;
    for (int k = 0; k < 6; ++k) st.stop_histogram[k] = hist[k];
    if (class_out) std::memcpy(class_out, cls.data(), n);
    if (stats) *stats = st;
    return 0;
}

// mask only, over an existing base buffer (used to check grrt_select_mask_device)
int orc_select_mask(const GrrtAdaptive* ad, const double* xyza, const uint8_t* cls, uint32_t H, uint32_t W,
                    uint32_t* selected, uint32_t* count, double* min_lum_out) {
    double min_lum = resolve_minimum_luminance(*ad, xyza, (size_t)H * W);
    static const int SH[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
    uint32_t n = 0;
    for (uint32_t row = 0; row < H; ++row)
        for (uint32_t col = 0; col < W; ++col) {
            size_t pi = (size_t)row * W + col;
            for (int k = 0; k < 8; ++k) {
                int nr = (int)row + SH[k][0], nc = (int)col + SH[k][1];
                if (nr < 0 || nr >= (int)H || nc < 0 || nc >= (int)W) continue;
                size_t ni = (size_t)nr * W + nc;
                if (should_supersample_pair(xyza + 4 * pi, cls[pi], xyza + 4 * ni, cls[ni], *ad, min_lum)) {
                    selected[n++] = (uint32_t)pi;
                    break;
                }
            }
        }
    *count = n;
    *min_lum_out = min_lum;
    return 0;
}

// xyz_to_linear_srgb_buffer + linear_srgb_to_srgb_buffer, color.rs:236-299
int orc_tonemap(const double* xyza, size_t n, int method, double exposure, uint8_t* rgb_out) {
    std::vector<double> lin(3 * n);
    for (size_t i = 0; i < n; ++i) xyz_to_linear_srgb(xyza + 4 * i, &lin[3 * i]);
    double scale = 1.0;
    if (method == GRRT_TONEMAP_GLOBAL_LINEAR) {
        double mr = 0.0, mg = 0.0, mb = 0.0;
        for (size_t i = 0; i < n; ++i) {
            mr = rmax(mr, lin[3 * i] * exposure);
            mg = rmax(mg, lin[3 * i + 1] * exposure);
            mb = rmax(mb, lin[3 * i + 2] * exposure);
        }
        double mc = rmax(rmax(mr, mg), mb);
        scale = mc > 0.0 ? 1.0 / mc : 1.0;
    }
    for (size_t i = 0; i < n; ++i) {
        double c[3] = {lin[3 * i] * exposure, lin[3 * i + 1] * exposure, lin[3 * i + 2] * exposure};
        if (method == GRRT_TONEMAP_REINHARD) {
            double l_in = 0.2126 * c[0] + 0.7152 * c[1] + 0.0722 * c[2];
            if (l_in > 0.0) {
                double l_out = l_in / (1.0 + l_in);
                double f = l_out / l_in;
                c[0] *= f; c[1] *= f; c[2] *= f;
            }
        } else {
            c[0] = scale * c[0]; c[1] = scale * c[1]; c[2] = scale * c[2];
        }
        for (int k = 0; k < 3; ++k) rgb_out[3 * i + k] = to_u8(compand_srgb(rmax(c[k], 0.0)) * 255.0);
    }
    return 0;
}
void orc_xyz_to_srgb(const double xyz[3], double exposure, uint8_t out[3]) {  // color.rs:220-234
    double v[3];
    xyz_to_linear_srgb(xyz, v);
    for (int k = 0; k < 3; ++k) out[k] = to_u8(compand_srgb(rmax(v[k] * exposure, 0.0)) * 255.0);
}

}  // extern "C"

// Same contract as grrt_trace_items_device (include/grrt.h), on the CPU: lets the multi-rank host logic of
// gr_raytracer_b200/renderer.py be tested under gloo without a GPU (tests/test_distributed_cpu.py).
extern "C" int orc_trace_items(const GrrtScene* s, const uint32_t* items, uint64_t first_item, uint64_t n_items,
                               uint32_t from_row, uint32_t from_col, uint32_t to_row, uint32_t to_col, uint32_t spa,
                               double* xyza, uint8_t* cls, uint8_t* stop, uint32_t* steps) {
    SceneView sv = view_of(s);
    (void)to_row;
    const uint32_t W = to_col - from_col;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < (int64_t)n_items; ++i) {
        uint64_t pid = items ? items[i] : first_item + (uint64_t)i;
        int64_t row = (int64_t)(pid / W) + from_row, col = (int64_t)(pid % W) + from_col;
        if (spa == 0) {
            RayResult r = trace_ray(sv, camera_ray(s->camera, row, col), true, [](const StepRec&) {});
            double* o = xyza + 4 * i;
            if (r.error) {
                o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 1.0;
                if (cls) cls[i] = GRRT_CLASS_ESCAPED;
            } else {
                o[0] = r.color.x; o[1] = r.color.y; o[2] = r.color.z; o[3] = r.color.alpha;
                if (cls) cls[i] = (uint8_t)r.ray_class;
            }
            if (stop) stop[i] = (uint8_t)r.stop;
            if (steps) steps[i] = (uint32_t)r.steps;
        } else {
            Cie acc = cie(0.0, 0.0, 0.0, 0.0);
            uint32_t valid = 0;
            for (uint64_t sr = 0; sr < spa; ++sr)
                for (uint64_t sc2 = 0; sc2 < spa; ++sc2) {
                    double dx, dy;
                    stratified_sample_offset(row, col, sr, sc2, spa, dx, dy);
                    RayResult r = trace_ray(sv, camera_ray_offset(s->camera, row, col, dx, dy), true, [](const StepRec&) {});
                    if (r.error) continue;
                    acc = cie(acc.x + r.color.x, acc.y + r.color.y, acc.z + r.color.z, acc.alpha + r.color.alpha);
                    ++valid;
                }
            if (valid > 0) {
                double inv = 1.0 / (double)valid;
                double* o = xyza + 4 * pid;
                o[0] = acc.x * inv; o[1] = acc.y * inv; o[2] = acc.z * inv; o[3] = acc.alpha * inv;
            }
        }
    }
    return 0;
}
