// scene_build.cpp -- host-side scene assembly: RenderConfig (TOML) + GlobalOpts -> GrrtScene.
//
// This is the C++ host code that sits above the C ABI where the reference has Rust:
//   src/configuration.rs:4-219      schema, defaults, AdaptiveSamplingConfig::validate
//   src/cli/shared.rs:23-86         resolve_camera_velocity, assert_future_directed
//   src/cli/shared.rs:131-350       create_scene, get_texture_mapper
//   src/cli/{schwarzschild,kerr,kerr_bl,euclidean,euclidean_spherical}.rs   chart of the camera position
//   src/rendering/camera.rs:27-196  Camera::new (tetrad, rotations, boost, handedness)
//   src/geometry/*.rs               tetrads, Lorentz boosts, observers
//   src/rendering/texture.rs:121-138, black_body_radiation.rs   black-body LUT
//   src/rendering/temperature.rs:45-195                          disc temperature LUT
// It runs once per frame.  It is compiled with -ffp-contract=off and keeps the reference's operation
// order, because the camera tetrad defines every ray and the Kerr-Schild integration is sensitive to the
// last bit of its initial condition (see csrc/grrt_device.cuh).
#include "scene_build.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "png.hpp"

namespace grrt_host {
namespace {

const double PI = 3.14159265358979323846;

using Vec4 = std::array<double, 4>;
using Mat4 = std::array<std::array<double, 4>, 4>;

Vec4 add(const Vec4& a, const Vec4& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}; }
Vec4 scale(double s, const Vec4& a) { return {s * a[0], s * a[1], s * a[2], s * a[3]}; }
Vec4 neg(const Vec4& a) { return {-a[0], -a[1], -a[2], -a[3]}; }
Vec4 matvec(const Mat4& m, const Vec4& v) {  // nalgebra gemv: k ascending
    Vec4 r;
    for (int i = 0; i < 4; ++i) r[i] = ((m[i][0] * v[0] + m[i][1] * v[1]) + m[i][2] * v[2]) + m[i][3] * v[3];
    return r;
}
double bilinear(const Vec4& v, const Mat4& g, const Vec4& w) {  // (v^T g) w
    double row[4];
    for (int j = 0; j < 4; ++j) row[j] = ((v[0] * g[0][j] + v[1] * g[1][j]) + v[2] * g[2][j]) + v[3] * g[3][j];
    return ((row[0] * w[0] + row[1] * w[1]) + row[2] * w[2]) + row[3] * w[3];
}
Mat4 zero4() {
    Mat4 m;
    for (auto& r : m) r.fill(0.0);
    return m;
}
double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
double rem_euclid(double a, double b) {
    double r = std::fmod(a, b);
    return r < 0.0 ? r + std::fabs(b) : r;
}

// ---- circular_orbit.rs -----------------------------------------------------------------------------
struct MetricTPhi {
    double tt, tphi, phiphi;
};
MetricTPhi bl_components_at(double r_s, double a, double r, double theta) {  // :39-46
    double c = std::cos(theta), s = std::sin(theta);
    double sig = r * r + a * a * (c * c);
    double sin2 = s * s;
    MetricTPhi m;
    m.tt = -(1.0 - r_s * r / sig);
    m.tphi = -a * r_s * r * sin2 / sig;
    m.phiphi = (r * r + a * a + a * a * r_s * r * sin2 / sig) * sin2;
    return m;
}
void zamo_coefficients(double r_s, double a, double r, double theta, double& u_t, double& u_phi) {  // :59-72
    MetricTPhi m = bl_components_at(r_s, a, r, theta);
    double omega = -m.tphi / m.phiphi;
    u_t = std::sqrt(-1.0 / (m.tt + 2.0 * m.tphi * omega + m.phiphi * omega * omega));
    u_phi = omega * u_t;
}
double orbit_angular_velocity(double r_s, double a, double r) {  // :76-80
    double m = 0.5 * r_s;
    double sm = std::sqrt(m);
    return sm / (std::pow(r, 1.5) + a * sm);
}
bool orbit_coefficients(double r_s, double a, double r, double& u_t, double& u_phi) {  // :84-107
    double omega = orbit_angular_velocity(r_s, a, r);
    MetricTPhi m = bl_components_at(r_s, a, r, PI / 2.0);
    double pre = m.tt + 2.0 * omega * m.tphi + omega * omega * m.phiphi;
    if (pre >= 0.0) return false;
    u_t = 1.0 / std::sqrt(-pre);
    u_phi = omega * u_t;
    return true;
}
bool orbit_energy(double r_s, double a, double r, double& out) {  // :110-115
    double omega = orbit_angular_velocity(r_s, a, r);
    MetricTPhi m = bl_components_at(r_s, a, r, PI / 2.0);
    double u_t, u_phi;
    if (!orbit_coefficients(r_s, a, r, u_t, u_phi)) return false;
    out = -(m.tt + m.tphi * omega) * u_t;
    return true;
}
bool orbit_angular_momentum(double r_s, double a, double r, double& out) {  // :118-123
    double omega = orbit_angular_velocity(r_s, a, r);
    MetricTPhi m = bl_components_at(r_s, a, r, PI / 2.0);
    double u_t, u_phi;
    if (!orbit_coefficients(r_s, a, r, u_t, u_phi)) return false;
    out = (m.tphi + m.phiphi * omega) * u_t;
    return true;
}
double isco_radius(double r_s, double a) {  // :126-136
    double a_s = 2.0 * a / r_s;
    double z1 = 1.0 + std::pow(1.0 - a_s * a_s, 1.0 / 3.0) * (std::pow(1.0 + a_s, 1.0 / 3.0) + std::pow(1.0 - a_s, 1.0 / 3.0));
    double z2 = std::sqrt(3.0 * a_s * a_s + z1 * z1);
    return (3.0 + z2 - std::sqrt((3.0 - z1) * (3.0 + z1 + 2.0 * z2))) * r_s / 2.0;
}

// ---- the geometries as seen by the camera construction ---------------------------------------------
enum Chart { CART, SPH, BL };

struct Frame {  // tetrad
    Vec4 t, x, y, z;
};

class HostGeometry {
public:
    GrrtGeometry g;
    explicit HostGeometry(const GrrtGeometry& gg) : g(gg) {}
    Chart chart() const {
        if (g.kind == GRRT_GEOM_EUCLIDEAN || g.kind == GRRT_GEOM_KERR) return CART;
        return g.kind == GRRT_GEOM_KERR_BL ? BL : SPH;
    }
    double sig0() const { return (g.kind == GRRT_GEOM_KERR || g.kind == GRRT_GEOM_KERR_BL) ? -1.0 : 1.0; }

    // --- Kerr-Schild pieces (kerr.rs:31-84)
    static double ks_r_sqr(double a, double x, double y, double z) {
        double rho = x * x + y * y + z * z;
        return 0.5 * (rho - a * a + std::sqrt((rho - a * a) * (rho - a * a) + 4.0 * a * a * z * z));
    }
    void ks_f_k(const Vec4& p, double& f, double k[4]) const {
        double x = p[1], y = p[2], z = p[3], a = g.a;
        double r2 = ks_r_sqr(a, x, y, z);
        double r = std::sqrt(r2);
        f = (r * r * r * g.radius) / (r * r * r * r + a * a * z * z);
        k[0] = 1.0;
        k[1] = (r * x + a * y) / (r2 + a * a);
        k[2] = (r * y - a * x) / (r2 + a * a);
        k[3] = z / r;
    }
    Mat4 ks_metric(const Vec4& p) const {
        double f, k[4];
        ks_f_k(p, f, k);
        Mat4 m = zero4();
        m[0][0] = k[0] * k[0] * f - 1.0;
        m[0][1] = k[0] * k[1] * f;
        m[0][2] = k[0] * k[2] * f;
        m[0][3] = k[0] * k[3] * f;
        m[1][1] = k[1] * k[1] * f + 1.0;
        m[1][2] = k[1] * k[2] * f;
        m[1][3] = k[1] * k[3] * f;
        m[2][2] = k[2] * k[2] * f + 1.0;
        m[2][3] = k[2] * k[3] * f;
        m[3][3] = k[3] * k[3] * f + 1.0;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < i; ++j) m[i][j] = m[j][i];
        return m;
    }
    // --- Boyer-Lindquist metric (kerr_bl.rs:253-272)
    Mat4 bl_metric(const Vec4& p) const {
        double r = p[1], theta = p[2], a = g.a, r_s = g.radius;
        double c = std::cos(theta), s = std::sin(theta);
        double sig = r * r + a * a * (c * c);
        double sin2 = s * s;
        Mat4 m = zero4();
        m[0][0] = -(1.0 - r_s * r / sig);
        m[1][1] = sig / (r * r - r_s * r + a * a);
        m[2][2] = sig;
        m[3][3] = (r * r + a * a + a * a * r_s * r * sin2 / sig) * sin2;
        m[0][3] = -a * r_s * r * sin2 / sig;
        m[3][0] = m[0][3];
        return m;
    }

    double inner(const Vec4& pos, const Vec4& v, const Vec4& w) const {
        switch (g.kind) {
            case GRRT_GEOM_EUCLIDEAN:  // euclidean.rs:61-67
                return 1.0 * v[0] * w[0] + -v[1] * w[1] + -v[2] * w[2] + -v[3] * w[3];
            case GRRT_GEOM_EUCLIDEAN_SPHERICAL: {  // euclidean_spherical.rs:80-90
                double r = pos[1], s = std::sin(pos[2]);
                return 1.0 * v[0] * w[0] - v[1] * w[1] - r * r * v[2] * w[2] - r * r * s * s * v[3] * w[3];
            }
            case GRRT_GEOM_SCHWARZSCHILD: {  // schwarzschild.rs:90-102
                double r = pos[1], s = std::sin(pos[2]);
                double a = 1.0 - g.radius / r;
                return a * v[0] * w[0] - v[1] * w[1] / a - r * r * v[2] * w[2] - r * r * s * s * v[3] * w[3];
            }
            case GRRT_GEOM_KERR: return bilinear(v, ks_metric(pos), w);  // kerr.rs:283-287
            default: {  // kerr_bl.rs:339-358
                Mat4 m = bl_metric(pos);
                double result = 0.0;
                for (int mu = 0; mu < 4; ++mu)
                    for (int nu = 0; nu < 4; ++nu) result += m[mu][nu] * v[mu] * w[nu];
                return result;
            }
        }
    }

    // main.rs:92-97 hands over a Cartesian point; each geometry converts it to its chart
    Vec4 camera_position_native(const double c[3]) const {
        double x = c[0], y = c[1], z = c[2];
        switch (chart()) {
            case CART: return {0.0, x, y, z};
            case SPH: {  // spherical_coordinates_helper.rs:5-26
                double r = std::sqrt(x * x + y * y + z * z);
                if (r == 0.0) return {0.0, 0.0, 0.0, 0.0};
                return {0.0, r, std::acos(z / r), std::atan2(y, x)};
            }
            default: {  // spherical_coordinates_helper.rs:44-61
                double a = g.a;
                double rho = x * x + y * y + z * z;
                double w = rho - a * a;
                double r2 = 0.5 * (w + std::sqrt(w * w + 4.0 * a * a * z * z));
                double r = std::sqrt(r2);
                double theta = r == 0.0 ? 0.0 : std::acos(clampd(z / r, -1.0, 1.0));
                double phi = std::atan2(r * y - a * x, r * x + a * y);
                return {0.0, r, theta, phi};
            }
        }
    }
    Vec4 axial_killing(const Vec4& pos) const {
        if (chart() == CART) return {0.0, -pos[2], pos[1], 0.0};
        return {0.0, 0.0, 0.0, 1.0};
    }
    Vec4 stationary_velocity(const Vec4& pos) const {
        switch (g.kind) {
            case GRRT_GEOM_SCHWARZSCHILD: return {1.0 / std::sqrt(1.0 - g.radius / pos[1]), 0.0, 0.0, 0.0};
            case GRRT_GEOM_KERR: {
                double f, k[4];
                ks_f_k(pos, f, k);
                return {1.0 / std::sqrt(1.0 - f), 0.0, 0.0, 0.0};
            }
            case GRRT_GEOM_KERR_BL: {
                double c = std::cos(pos[2]);
                double sig = pos[1] * pos[1] + g.a * g.a * (c * c);
                return {1.0 / std::sqrt(1.0 - g.radius * pos[1] / sig), 0.0, 0.0, 0.0};
            }
            default: return {1.0, 0.0, 0.0, 0.0};
        }
    }
    Vec4 zamo_velocity(const Vec4& pos) const {
        double u_t, u_phi;
        if (g.kind == GRRT_GEOM_KERR) {  // kerr.rs:457-468
            double r = std::sqrt(ks_r_sqr(g.a, pos[1], pos[2], pos[3]));
            double theta = r == 0.0 ? 0.0 : std::acos(clampd(pos[3] / r, -1.0, 1.0));
            zamo_coefficients(g.radius, g.a, r, theta, u_t, u_phi);
            return add(scale(u_t, Vec4{1.0, 0.0, 0.0, 0.0}), scale(u_phi, axial_killing(pos)));
        }
        if (g.kind == GRRT_GEOM_KERR_BL) {  // kerr_bl.rs:373-382
            zamo_coefficients(g.radius, g.a, pos[1], pos[2], u_t, u_phi);
            return {u_t, 0.0, 0.0, u_phi};
        }
        return stationary_velocity(pos);
    }

    // gram_schmidt.rs:6-34 (modified Gram-Schmidt)
    void gram_schmidt(const Vec4& pos, const Vec4 in[4], Vec4 out[4]) const {
        for (int i = 0; i < 4; ++i) {
            Vec4 w = in[i];
            for (int j = 0; j < i; ++j) {
                double p1 = inner(pos, w, out[j]);
                double p2 = inner(pos, out[j], out[j]);
                w = add(w, neg(scale(p1 / p2, out[j])));
            }
            double n = std::sqrt(std::fabs(inner(pos, w, w)));
            out[i] = scale(1.0 / n, w);
        }
    }

    Frame tetrad_at(const Vec4& pos) const {
        Frame F;
        switch (g.kind) {
            case GRRT_GEOM_EUCLIDEAN: {  // euclidean.rs:88-110
                double r = std::sqrt(pos[1] * pos[1] + pos[2] * pos[2] + pos[3] * pos[3]);
                double theta = 0.0, phi = 0.0;
                if (r != 0.0) {
                    theta = std::acos(pos[3] / r);
                    phi = std::atan2(pos[2], pos[1]);
                }
                Vec4 e_r = {0.0, std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta)};
                Vec4 e_th = {0.0, std::cos(theta) * std::cos(phi), std::cos(theta) * std::sin(phi), -std::sin(theta)};
                Vec4 e_ph = {0.0, -std::sin(phi), std::cos(phi), 0.0};
                F.t = {1.0, 0.0, 0.0, 0.0};
                F.x = e_ph;
                F.y = neg(e_th);
                F.z = neg(e_r);
                return F;
            }
            case GRRT_GEOM_EUCLIDEAN_SPHERICAL: {  // euclidean_spherical.rs:100-111
                double r = pos[1], theta = pos[2];
                F.t = {1.0, 0.0, 0.0, 0.0};
                F.x = {0.0, 0.0, 0.0, 1.0 / (r * std::sin(theta))};
                F.y = neg(Vec4{0.0, 0.0, 1.0 / r, 0.0});
                F.z = neg(Vec4{0.0, 1.0, 0.0, 0.0});
                return F;
            }
            case GRRT_GEOM_SCHWARZSCHILD: {  // schwarzschild.rs:116-132 (free-fall frame)
                double r = pos[1], theta = pos[2];
                double rr0 = g.radius / r;
                double a = 1.0 - rr0;
                F.t = {1.0 / a, -std::sqrt(rr0), 0.0, 0.0};
                F.x = {0.0, 0.0, 0.0, 1.0 / (r * std::sin(theta))};
                F.y = {0.0, 0.0, 1.0 / r, 0.0};
                F.z = {-std::sqrt(rr0) / a, 1.0, 0.0, 0.0};
                return F;
            }
            case GRRT_GEOM_KERR: {  // kerr.rs:297-331
                double f, k[4];
                ks_f_k(pos, f, k);
                double alpha = 1.0 / std::sqrt(1.0 + f);
                double bfac = f / (1.0 + f);
                double beta[3] = {bfac * k[1], bfac * k[2], bfac * k[3]};
                Vec4 in[4] = {{1.0 / alpha, -beta[0] / alpha, -beta[1] / alpha, -beta[2] / alpha},
                              {0.0, 1.0, 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0}, {0.0, 0.0, 0.0, 1.0}};
                Vec4 out[4];
                gram_schmidt(pos, in, out);
                F.t = out[0]; F.x = out[1]; F.y = out[2]; F.z = out[3];
                return F;
            }
            default: {  // kerr_bl.rs:428-450: ZAMO, then (phi, theta, r)
                double u_t, u_phi;
                zamo_coefficients(g.radius, g.a, pos[1], pos[2], u_t, u_phi);
                double omega = u_phi / u_t;
                Vec4 in[4] = {{u_t, 0.0, 0.0, u_t * omega}, {0.0, 0.0, 0.0, 1.0}, {0.0, 0.0, 1.0, 0.0}, {0.0, 1.0, 0.0, 0.0}};
                Vec4 out[4];
                gram_schmidt(pos, in, out);
                F.t = out[0]; F.x = out[1]; F.y = out[2]; F.z = out[3];
                return F;
            }
        }
    }

    Mat4 lorentz(const Vec4& pos, const Vec4& vel) const {
        Mat4 L = zero4();
        Vec4 tt = tetrad_at(pos).t;
        switch (g.kind) {
            case GRRT_GEOM_EUCLIDEAN: {  // euclidean.rs:112-142
                double gamma = tt[0] * vel[0] - tt[1] * vel[1] - tt[2] * vel[2] - tt[3] * vel[3];
                for (int i = 0; i < 4; ++i)
                    for (int j = 0; j < 4; ++j) {
                        double res = i == j ? 1.0 : 0.0;
                        double gg = j == 0 ? 1.0 : -1.0;
                        double a = -1.0 / (1.0 + gamma);
                        double b = tt[i] + vel[i];
                        double c = gg * (tt[j] + vel[j]);
                        res += a * b * c;
                        res += 2.0 * gg * tt[i] * vel[j];
                        L[i][j] = res;
                    }
                return L;
            }
            case GRRT_GEOM_EUCLIDEAN_SPHERICAL:  // euclidean_spherical.rs:113-121
                for (int i = 0; i < 4; ++i) L[i][i] = 1.0;
                return L;
            case GRRT_GEOM_SCHWARZSCHILD: {  // schwarzschild.rs:134-179
                double r = pos[1], s = std::sin(pos[2]);
                double a = 1.0 - g.radius / r;
                double md[4] = {a, -1.0 / a, -r * r, -r * r * s * s};
                double gamma = 0.0;
                for (int i = 0; i < 4; ++i) gamma += md[i] * vel[i] * tt[i];
                for (int mu = 0; mu < 4; ++mu)
                    for (int nu = 0; nu < 4; ++nu) {
                        double res = mu == nu ? 1.0 : 0.0;
                        double aa = 1.0 / (1.0 + gamma);
                        double b = tt[mu] + vel[mu];
                        double c = md[nu] * (tt[nu] + vel[nu]);
                        res -= aa * b * c;
                        res += 2.0 * md[nu] * tt[nu] * vel[mu];
                        L[mu][nu] = res;
                    }
                return L;
            }
            default: {  // kerr.rs:333-380, kerr_bl.rs:452-480
                Mat4 m = g.kind == GRRT_GEOM_KERR ? ks_metric(pos) : bl_metric(pos);
                double gamma = -bilinear(tt, m, vel);
                Vec4 uv = add(tt, vel);
                Vec4 uv_lower = matvec(m, uv);
                Vec4 mt = matvec(m, tt);
                for (int mu = 0; mu < 4; ++mu)
                    for (int nu = 0; nu < 4; ++nu) {
                        double res = mu == nu ? 1.0 : 0.0;
                        double a = 1.0 / (1.0 + gamma);
                        res += a * uv[mu] * uv_lower[nu];
                        res -= 2.0 * mt[nu] * vel[mu];
                        L[mu][nu] = res;
                    }
                return L;
            }
        }
    }

    bool orthonormal(const Vec4& pos, const Frame& F) const {  // tetrad.rs:60-130
        const double tol = 1e-5;
        double s0 = sig0(), s1 = -s0;
        auto near = [&](double a, double b) { return std::fabs(a - b) < tol; };
        return near(inner(pos, F.t, F.t), s0) && near(inner(pos, F.x, F.x), s1) && near(inner(pos, F.y, F.y), s1) &&
               near(inner(pos, F.z, F.z), s1) && near(inner(pos, F.t, F.x), 0.0) && near(inner(pos, F.t, F.y), 0.0) &&
               near(inner(pos, F.t, F.z), 0.0) && near(inner(pos, F.x, F.y), 0.0) && near(inner(pos, F.x, F.z), 0.0) &&
               near(inner(pos, F.y, F.z), 0.0);
    }

    // camera.rs:83-148
    void spatial_cartesian(const Vec4& pos, const Vec4& v, double out[3]) const {
        Chart c = chart();
        if (c == CART) {
            out[0] = v[1]; out[1] = v[2]; out[2] = v[3];
            return;
        }
        double r = pos[1], theta = rem_euclid(pos[2], PI), phi = rem_euclid(pos[3] + PI, 2.0 * PI) - PI;  // get_as_spherical
        double dr = v[1], dth = v[2], dph = v[3];
        double st = std::sin(theta), ct = std::cos(theta), sp = std::sin(phi), cp = std::cos(phi);
        if (c == SPH) {
            out[0] = st * cp * dr + r * ct * cp * dth - r * st * sp * dph;
            out[1] = st * sp * dr + r * ct * sp * dth + r * st * cp * dph;
            out[2] = ct * dr - r * st * dth;
        } else {
            double a = g.a;
            out[0] = st * cp * dr + (r * cp - a * sp) * ct * dth + (-r * sp - a * cp) * st * dph;
            out[1] = st * sp * dr + (r * sp + a * cp) * ct * dth + (r * cp - a * sp) * st * dph;
            out[2] = ct * dr - r * st * dth;
        }
    }
    double handedness(const Vec4& pos, const Frame& F) const {
        double x[3], y[3], z[3];
        spatial_cartesian(pos, F.x, x);
        spatial_cartesian(pos, F.y, y);
        spatial_cartesian(pos, F.z, z);
        double cr[3] = {y[1] * z[2] - y[2] * z[1], y[2] * z[0] - y[0] * z[2], y[0] * z[1] - y[1] * z[0]};
        double triple = (x[0] * cr[0] + x[1] * cr[1]) + x[2] * cr[2];
        if (!std::isfinite(triple) || std::fabs(triple) <= 1e-12) return chart() == CART ? 1.0 : -1.0;
        return triple >= 0.0 ? 1.0 : -1.0;
    }
};

void rotate_pair(const Vec4& v1, const Vec4& v2, double angle, Vec4& o1, Vec4& o2) {  // camera.rs:76-81
    o1 = add(scale(std::cos(angle), v1), scale(std::sin(angle), v2));
    o2 = add(scale(-std::sin(angle), v1), scale(std::cos(angle), v2));
}

// ---- colour helpers (color.rs, black_body_radiation.rs) -----------------------------------------------
double inv_compand(double u) { return u <= 0.04045 ? u / 12.92 : std::pow((u + 0.055) / 1.055, 2.4); }
void srgb8_to_xyz(const uint8_t c[3], double out[4]) {  // color.rs:310-332
    double r = inv_compand((double)c[0] / 255.0), g = inv_compand((double)c[1] / 255.0), b = inv_compand((double)c[2] / 255.0);
    out[0] = (0.4124564 * r + 0.3575761 * g) + 0.1804375 * b;
    out[1] = (0.2126729 * r + 0.7151522 * g) + 0.0721750 * b;
    out[2] = (0.0193339 * r + 0.1191920 * g) + 0.9503041 * b;
    out[3] = 1.0;
}
double cmf_lobe(double lambda, double mu, double tl, double tr) {  // color.rs:173-177
    double tau = lambda < mu ? tl : tr;
    double t = (lambda - mu) * tau;
    return std::exp(-0.5 * t * t);
}
void blackbody_xyz(double temperature, double redshift, double out[3]) {  // black_body_radiation.rs:12-41
    const double H = 6.62607015e-34, C = 299792458.0, KB = 1.380649e-23, NM = 1e-9;
    double interval = (830.0 - 380.0) * NM;
    double step = 1.0 * NM;
    double nsteps_f = std::floor(interval / step);
    unsigned long nsteps = nsteps_f > 0.0 ? (unsigned long)nsteps_f : 0;
    double xa = 0.0, ya = 0.0, za = 0.0;
    for (unsigned long i = 0; i < nsteps; ++i) {
        double lambda = 380.0 * NM + ((double)i + 0.5) * step;
        double l = lambda * redshift;
        double a = 2.0 * H * C * C;
        double b = H * C / (l * KB * temperature);
        double l2 = l * l;
        double radiance = a / (((l2 * l2) * l) * (std::exp(b) - 1.0));
        double nm = lambda / NM;
        double xb = 1.056 * cmf_lobe(nm, 599.8, 0.0264, 0.0323) + 0.362 * cmf_lobe(nm, 442.0, 0.0624, 0.0374) -
                    0.065 * cmf_lobe(nm, 501.1, 0.0490, 0.0382);
        double yb = 0.821 * cmf_lobe(nm, 568.8, 0.0213, 0.0247) + 0.286 * cmf_lobe(nm, 530.9, 0.0613, 0.0322);
        double zb = 1.217 * cmf_lobe(nm, 437.0, 0.0845, 0.0278) + 0.681 * cmf_lobe(nm, 459.0, 0.0385, 0.0725);
        xa += radiance * xb * step;
        ya += radiance * yb * step;
        za += radiance * zb * step;
    }
    double z2 = redshift * redshift;
    double boost = (z2 * z2) * redshift;
    out[0] = xa * boost;
    out[1] = ya * boost;
    out[2] = za * boost;
}

// ---- KerrTemperatureComputer::new (temperature.rs:45-195) ------------------------------------------------
struct DiscFlux {
    double a, radius, r_isco, m_dot;
    bool dl_dr(double r, double& o) const {
        double h = 1e-6 * std::fmax(r, 1.0);
        double lp, l0;
        if (r - h < r_isco) {
            if (!orbit_angular_momentum(radius, a, r + h, lp) || !orbit_angular_momentum(radius, a, r, l0)) return false;
            o = (lp - l0) / h;
        } else {
            if (!orbit_angular_momentum(radius, a, r + h, lp) || !orbit_angular_momentum(radius, a, r - h, l0)) return false;
            o = (lp - l0) / (2.0 * h);
        }
        return true;
    }
    bool integral(double r, double& o) const {
        const int N = 1000;
        double dr = (r - r_isco) / (double)N;
        double acc = 0.0;
        for (int i = 0; i < N; ++i) {
            double rp = r_isco + ((double)i + 0.5) * dr;
            double e, l, dl;
            if (!orbit_energy(radius, a, rp, e) || !orbit_angular_momentum(radius, a, rp, l)) return false;
            double om = orbit_angular_velocity(radius, a, rp);
            if (!dl_dr(rp, dl)) return false;
            acc += (e - om * l) * dl * dr;
        }
        o = acc;
        return true;
    }
    bool prefactor(double r, double& o) const {
        double e, l;
        if (!orbit_energy(radius, a, r, e) || !orbit_angular_momentum(radius, a, r, l)) return false;
        double om = orbit_angular_velocity(radius, a, r);
        double d = e - om * l;
        double denominator = (r * r) * (d * d);
        if (std::fabs(denominator) < 1e-20) return false;
        const double h = 1e-10;
        double d_omega = (orbit_angular_velocity(radius, a, r + h) - orbit_angular_velocity(radius, a, r - h)) / (2.0 * h);
        o = d_omega / denominator;
        return true;
    }
    bool flux(double r, double& o) const {
        double in, pre;
        if (!integral(r, in) || !prefactor(r, pre)) return false;
        double coefficient = -m_dot / (PI * radius * radius);
        o = coefficient * pre * in;
        return true;
    }
};

bool build_temperature_lut(double temperature, double outer_radius, double a_in, double radius, std::vector<double>& lut_r,
                           std::vector<double>& lut_t, double& r_isco_out) {
    const int N = GRRT_TEMPERATURE_LUT_N;
    double a = std::fabs(a_in);
    double r_isco = isco_radius(radius, a);
    double outer = outer_radius <= r_isco ? r_isco + std::fmax(1e-6, std::fabs(r_isco) * 1e-9) : outer_radius;
    DiscFlux d{a, radius, r_isco, 1.0};
    double max_f = 0.0, max_r = 0.0;
    double dr = (outer - r_isco) / 10.0;
    for (int i = 0; i < 10; ++i) {
        double r = r_isco + ((double)i + 0.5) * dr;
        double f;
        if (!d.flux(r, f)) return false;
        if (max_f < f) {
            max_f = f;
            max_r = r;
        }
    }
    double in, pre;
    if (!d.integral(max_r, in) || !d.prefactor(max_r, pre)) return false;
    double coefficient = -1.0 / (PI * radius * radius);
    double target = 1.0 * std::pow(temperature, 4.0);
    d.m_dot = target / (coefficient * pre * in);
    lut_r.resize(N);
    lut_t.resize(N);
    double step = (outer - r_isco) / (double)(N - 1);
    bool ok = true;
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < N; ++i) {
        double r = r_isco + (double)i * step;
        double f = 0.0;
        if (!d.flux(r, f)) ok = false;
        lut_r[i] = r;
        lut_t[i] = std::pow(std::fmax(f / 1.0, 0.0), 0.25);
    }
    r_isco_out = r_isco;
    return ok;
}

// ---- Perlin permutation table of noise 0.9.0 (PermutationTable::new; rand_xorshift 0.3 + rand 0.8 shuffle) -----
void perlin_table(uint32_t seed, uint8_t out[256]) {
    uint32_t s[4] = {1u, seed, seed, seed};  // bytes: real[0] = 1, real[4..16] = seed (little endian) x 3
    if (!s[0] && !s[1] && !s[2] && !s[3]) s[0] = s[1] = s[2] = s[3] = 0x0BAD5EEDu;
    uint32_t x = s[0], y = s[1], z = s[2], w = s[3];
    auto next = [&]() {
        uint32_t t = x ^ (x << 11);
        x = y;
        y = z;
        z = w;
        w = w ^ (w >> 19) ^ (t ^ (t >> 8));
        return w;
    };
    for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
    for (uint32_t i = 255; i >= 1; --i) {
        uint32_t range = i + 1;
        uint32_t zone = (range << __builtin_clz(range)) - 1u;
        uint32_t j;
        for (;;) {
            uint64_t m = (uint64_t)next() * range;
            if ((uint32_t)m <= zone) {
                j = (uint32_t)(m >> 32);
                break;
            }
        }
        std::swap(out[i], out[j]);
    }
}

// ---- TOML accessors ---------------------------------------------------------------------------------------
struct ConfigError {
    int code;
    std::string msg;
};
[[noreturn]] void bad(int code, const std::string& msg) { throw ConfigError{code, msg}; }

const TomlValue& need(const TomlValue& t, const char* key, const char* where) {
    const TomlValue* v = t.find(key);
    if (!v) bad(GRRT_ERR_TOML, std::string("missing field `") + key + "` in " + where);
    return *v;
}
double need_f64(const TomlValue& t, const char* key, const char* where) {
    const TomlValue& v = need(t, key, where);
    if (!v.is_number()) bad(GRRT_ERR_TOML, std::string("field `") + key + "` in " + where + " must be a number");
    return v.as_double();
}
uint64_t need_usize(const TomlValue& t, const char* key, const char* where) {
    const TomlValue& v = need(t, key, where);
    if (v.kind != TomlValue::INT || v.i < 0) bad(GRRT_ERR_TOML, std::string("field `") + key + "` in " + where + " must be a non-negative integer");
    return (uint64_t)v.i;
}
void need_vec3(const TomlValue& t, const char* key, const char* where, double out[3]) {
    const TomlValue& v = need(t, key, where);
    if (v.kind != TomlValue::ARRAY || v.arr.size() != 3) bad(GRRT_ERR_TOML, std::string("field `") + key + "` in " + where + " must have 3 elements");
    for (int i = 0; i < 3; ++i) {
        if (!v.arr[i].is_number()) bad(GRRT_ERR_TOML, std::string("field `") + key + "` must hold numbers");
        out[i] = v.arr[i].as_double();
    }
}
// serde externally tagged enum: "Variant" or { Variant = { ... } }
void enum_variant(const TomlValue& v, const char* what, std::string& name, const TomlValue*& body) {
    static const TomlValue empty = [] {
        TomlValue t;
        t.kind = TomlValue::TABLE;
        return t;
    }();
    if (v.kind == TomlValue::STRING) {
        name = v.s;
        body = &empty;
        return;
    }
    if (v.kind == TomlValue::TABLE && v.tab.size() == 1) {
        name = v.tab.begin()->first;
        body = &v.tab.begin()->second;
        if (body->kind != TomlValue::TABLE) bad(GRRT_ERR_TOML, std::string("variant of ") + what + " must be a table");
        return;
    }
    bad(GRRT_ERR_TOML, std::string("expected exactly one variant for ") + what);
}

}  // namespace

// get_cie_xyz_of_black_body_redshifted (black_body_radiation.rs:12-41) for the `blackbody` utilities
void blackbody_xyz_redshifted(double temperature, double redshift, double out[3]) { blackbody_xyz(temperature, redshift, out); }

// xyz_to_srgb (color.rs:218-233): IEC 2003 matrix, exposure, clamp, sRGB companding, round -- no tone mapping
void xyz_to_srgb8(const double xyz[3], double exposure, uint8_t rgb[3]) {
    const double m[3][3] = {{3.2406255, -1.5372080, -0.4986286}, {-0.9689307, 1.8757561, 0.0415175}, {0.0557101, -0.2040211, 1.0569959}};
    for (int k = 0; k < 3; ++k) {
        double lin = ((m[k][0] * xyz[0] + m[k][1] * xyz[1]) + m[k][2] * xyz[2]) * exposure;
        lin = lin > 0.0 ? lin : 0.0;
        double enc = lin <= 0.0031308 ? 12.92 * lin : 1.055 * std::pow(lin, 1.0 / 2.4) - 0.055;
        enc = enc < 0.0 ? 0.0 : (enc > 1.0 ? 1.0 : enc);
        rgb[k] = (uint8_t)std::round(enc * 255.0);
    }
}


// =========================================================================================================
// resolve_camera_velocity + assert_future_directed (cli/shared.rs:23-86)
Vec4 resolve_camera_velocity(const HostGeometry& geo, const Vec4& pos, const std::string& mode, const double explicit_components[4]) {
    Vec4 vel;
    if (mode == "StaticObserver") vel = geo.stationary_velocity(pos);
    else if (mode == "Zamo") vel = geo.zamo_velocity(pos);
    else if (mode == "Explicit") {
        for (int i = 0; i < 4; ++i) vel[i] = explicit_components[i];
        double n = geo.inner(pos, vel, vel);
        if (std::fabs(n - geo.sig0()) > 1e-6) bad(GRRT_ERR_INVALID_CONFIGURATION, "Explicit camera_velocity is not normalized");
    } else {
        bad(GRRT_ERR_TOML, "unknown variant `" + mode + "` for camera_velocity");
    }
    Vec4 tt = geo.tetrad_at(pos).t;
    double orientation = geo.sig0() * geo.inner(pos, tt, vel);
    if (!(std::isfinite(vel[0]) && 1.0 * orientation > 0.0))
        bad(GRRT_ERR_INVALID_CONFIGURATION, "camera four-velocity has wrong time orientation");
    return vel;
}

// Camera::new (camera.rs:151-196): alpha = PI/4, rows = height, cols = width
void make_camera(const HostGeometry& geo, const Vec4& pos, const Vec4& vel, const GrrtCliOpts& opts, GrrtCamera& cam) {
    Frame T0 = geo.tetrad_at(pos);
    if (!geo.orthonormal(pos, T0)) bad(GRRT_ERR_TETRAD_NOT_ORTHONORMAL, "Tetrad is not orthonormal");
    Vec4 a1, b1, z, a2, x, y;
    rotate_pair(T0.x, T0.y, opts.phi, a1, b1);
    rotate_pair(T0.z, a1, opts.theta, z, a2);
    rotate_pair(a2, b1, opts.psi, x, y);
    Mat4 L = geo.lorentz(pos, vel);
    Frame T{matvec(L, T0.t), matvec(L, x), matvec(L, y), matvec(L, z)};
    if (!geo.orthonormal(pos, T)) bad(GRRT_ERR_TETRAD_NOT_ORTHONORMAL, "Tetrad is not orthonormal");
    for (int i = 0; i < 4; ++i) {
        cam.position[i] = pos[i];
        cam.velocity[i] = vel[i];
        cam.tetrad[0][i] = T.t[i];
        cam.tetrad[1][i] = T.x[i];
        cam.tetrad[2][i] = T.y[i];
        cam.tetrad[3][i] = T.z[i];
    }
    cam.alpha = PI / 4.0;
    cam.rows = opts.height;
    cam.cols = opts.width;
    cam.spatial_signature = -geo.sig0();
    cam.spatial_handedness = geo.handedness(pos, T);
}

struct SceneBuilder {
    HostSceneData& hs;
    const GrrtCliOpts& opts;
    std::string resource_root;
    std::map<std::string, std::pair<double, size_t>> bitmap_cache;  // path -> (first beaming exponent, index into hs.bitmaps)

    SceneBuilder(HostSceneData& h, const GrrtCliOpts& o, const char* root) : hs(h), opts(o), resource_root(root ? root : "") {}

    void ensure_blackbody_lut() {  // BlackBodyMapper::new, texture.rs:121-138
        if (!hs.bb_log_t.empty()) return;
        const int N = GRRT_BLACKBODY_LUT_N;
        hs.bb_log_t.resize(N);
        hs.bb_xyz.resize(3 * (size_t)N);
        double min_log = std::log10(10.0), max_log = std::log10(10000000.0);
        double step = (max_log - min_log) / (double)(N - 1);
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = 0; i < N; ++i) {
            double lt = min_log + (double)i * step;
            double t = std::pow(10.0, lt);
            hs.bb_log_t[i] = lt;
            blackbody_xyz(t, 1.0, &hs.bb_xyz[3 * (size_t)i]);
        }
    }

    GrrtTexture texture(const TomlValue& cfg) {  // get_texture_mapper, cli/shared.rs:323-350
        GrrtTexture t;
        std::memset(&t, 0, sizeof t);
        std::string name;
        const TomlValue* b;
        enum_variant(cfg, "texture", name, b);
        if (name == "Bitmap") {
            double beaming = need_f64(*b, "beaming_exponent", "Bitmap");
            const TomlValue& pv = need(*b, "path", "Bitmap");
            if (pv.kind != TomlValue::STRING) bad(GRRT_ERR_TOML, "Bitmap.path must be a string");
            auto it = bitmap_cache.find(pv.s);
            if (it == bitmap_cache.end()) {  // TextureMapperFactory (texture.rs:260-296): one mapper per file name
                std::string full = pv.s;
                if (!resource_root.empty() && !(full.size() && full[0] == '/')) full = resource_root + "/" + full;
                HostBitmap bm;
                std::string err;
                if (!decode_png_rgba8(full, bm.rgba, bm.width, bm.height, err)) bad(GRRT_ERR_TEXTURE, "Texture error: " + err);
                hs.bitmaps.push_back(std::move(bm));
                it = bitmap_cache.emplace(pv.s, std::make_pair(beaming, hs.bitmaps.size() - 1)).first;
            }
            t.kind = GRRT_TEX_BITMAP;
            t.beaming_exponent = it->second.first;  // the first user's exponent sticks to the cached mapper
            t.width = hs.bitmaps[it->second.second].width;
            t.height = hs.bitmaps[it->second.second].height;
            t._pad = (uint32_t)it->second.second;   // resolved to a pointer in finish()
        } else if (name == "Checker") {
            t.kind = GRRT_TEX_CHECKER;
            t.beaming_exponent = need_f64(*b, "beaming_exponent", "Checker");
            t.checker_width = need_f64(*b, "width", "Checker");
            t.checker_height = need_f64(*b, "height", "Checker");
            const char* keys[2] = {"color1", "color2"};
            double* dst[2] = {t.c1, t.c2};
            for (int k = 0; k < 2; ++k) {
                const TomlValue& cv = need(*b, keys[k], "Checker");
                if (cv.kind != TomlValue::ARRAY || cv.arr.size() != 3) bad(GRRT_ERR_TOML, "Checker colours need 3 components");
                uint8_t c[3];
                for (int i = 0; i < 3; ++i) {
                    if (cv.arr[i].kind != TomlValue::INT || cv.arr[i].i < 0 || cv.arr[i].i > 255) bad(GRRT_ERR_TOML, "Checker colour components must be 0..255");
                    c[i] = (uint8_t)cv.arr[i].i;
                }
                srgb8_to_xyz(c, dst[k]);
            }
        } else if (name == "BlackBody") {
            t.kind = GRRT_TEX_BLACKBODY;
            t.beaming_exponent = need_f64(*b, "beaming_exponent", "BlackBody");
            ensure_blackbody_lut();
        } else {
            bad(GRRT_ERR_TOML, "unknown variant `" + name + "` for texture");
        }
        return t;
    }

    GrrtTemperature temperature(const GrrtGeometry& g, double temperature, double outer_radius) {
        GrrtTemperature tc;
        std::memset(&tc, 0, sizeof tc);
        if (g.kind == GRRT_GEOM_EUCLIDEAN || g.kind == GRRT_GEOM_EUCLIDEAN_SPHERICAL) {  // euclidean.rs:218-225
            tc.kind = GRRT_TEMP_CONSTANT;
            tc.constant = temperature;
            return tc;
        }
        double a = g.kind == GRRT_GEOM_SCHWARZSCHILD ? 0.0 : g.a;  // schwarzschild.rs:267-279
        hs.temp_luts.emplace_back();
        HostTempLut& lut = hs.temp_luts.back();
        if (!build_temperature_lut(temperature, outer_radius, a, g.radius, lut.r, lut.t, lut.r_isco))
            bad(GRRT_ERR_LUT, "temperature LUT construction failed (no circular orbit / denominator close to zero)");
        tc.kind = GRRT_TEMP_LUT;
        tc.n = GRRT_TEMPERATURE_LUT_N;
        tc.r_isco = lut.r_isco;
        tc.constant = (double)(hs.temp_luts.size() - 1);  // index, resolved to pointers in finish()
        return tc;
    }

    void adaptive(const TomlValue* cfg) {  // configuration.rs:21-93
        GrrtAdaptive& ad = hs.scene.adaptive;
        std::memset(&ad, 0, sizeof ad);
        ad.enabled = 1;
        ad.samples_per_axis = 4;
        ad.luminance_contrast_threshold = 0.15;
        ad.opacity_contrast_threshold = 0.1;
        ad.has_minimum_luminance = 0;
        ad.object_hit_opacity_threshold = 0.5;
        ad.exclude_background_contrast = 1;
        if (cfg) {
            if (cfg->kind != TomlValue::TABLE) bad(GRRT_ERR_TOML, "adaptive_sampling must be a table");
            for (const auto& kv : cfg->tab) {
                const TomlValue& v = kv.second;
                if (kv.first == "enabled" && v.kind == TomlValue::BOOL) ad.enabled = v.b;
                else if (kv.first == "samples_per_axis" && v.kind == TomlValue::INT && v.i >= 0) ad.samples_per_axis = (uint64_t)v.i;
                else if (kv.first == "luminance_contrast_threshold" && v.is_number()) ad.luminance_contrast_threshold = v.as_double();
                else if (kv.first == "opacity_contrast_threshold" && v.is_number()) ad.opacity_contrast_threshold = v.as_double();
                else if (kv.first == "minimum_luminance" && v.is_number()) { ad.has_minimum_luminance = 1; ad.minimum_luminance = v.as_double(); }
                else if (kv.first == "object_hit_opacity_threshold" && v.is_number()) ad.object_hit_opacity_threshold = v.as_double();
                else if (kv.first == "exclude_background_contrast" && v.kind == TomlValue::BOOL) ad.exclude_background_contrast = v.b;
                else bad(GRRT_ERR_TOML, "invalid field `" + kv.first + "` in adaptive_sampling");
            }
        }
        if (ad.samples_per_axis == 0) bad(GRRT_ERR_INVALID_CONFIGURATION, "adaptive_sampling.samples_per_axis must be greater than zero");
        const char* names[3] = {"luminance_contrast_threshold", "opacity_contrast_threshold", "object_hit_opacity_threshold"};
        double vals[3] = {ad.luminance_contrast_threshold, ad.opacity_contrast_threshold, ad.object_hit_opacity_threshold};
        for (int i = 0; i < 3; ++i)
            if (!std::isfinite(vals[i]) || !(vals[i] >= 0.0 && vals[i] <= 1.0))
                bad(GRRT_ERR_INVALID_CONFIGURATION, std::string("adaptive_sampling.") + names[i] + " must be finite and between 0 and 1");
        if (ad.has_minimum_luminance && (!std::isfinite(ad.minimum_luminance) || ad.minimum_luminance < 0.0))
            bad(GRRT_ERR_INVALID_CONFIGURATION, "adaptive_sampling.minimum_luminance must be finite and non-negative");
        ad.show_sampling_mask = opts.show_sampling_mask ? 1 : 0;
        if (opts.show_sampling_mask) {  // CIETristimulus::from_color (color.rs:36-40)
            srgb8_to_xyz(opts.sampling_mask_color, ad.mask_xyza);
            ad.mask_xyza[3] = 255.0 / 255.0;
        }
    }

    void build(const TomlValue& cfg) {
        GrrtScene& sc = hs.scene;
        std::memset(&sc, 0, sizeof sc);
        // geometry_type (configuration.rs:110-159); the serde default is Euclidean
        GrrtGeometry g;
        std::memset(&g, 0, sizeof g);
        {
            const TomlValue& gv = need(cfg, "geometry_type", "RenderConfig");
            std::string name;
            const TomlValue* b;
            enum_variant(gv, "geometry_type", name, b);
            if (name == "Euclidean") g.kind = GRRT_GEOM_EUCLIDEAN;
            else if (name == "EuclideanSpherical") g.kind = GRRT_GEOM_EUCLIDEAN_SPHERICAL;
            else if (name == "Schwarzschild") {
                g.kind = GRRT_GEOM_SCHWARZSCHILD;
                g.radius = need_f64(*b, "radius", "Schwarzschild");
                g.horizon_epsilon = need_f64(*b, "horizon_epsilon", "Schwarzschild");
            } else if (name == "Kerr" || name == "KerrBL") {
                g.kind = name == "Kerr" ? GRRT_GEOM_KERR : GRRT_GEOM_KERR_BL;
                g.radius = need_f64(*b, "radius", name.c_str());
                g.a = need_f64(*b, "a", name.c_str());
                g.horizon_epsilon = need_f64(*b, "horizon_epsilon", name.c_str());
            } else {
                bad(GRRT_ERR_TOML, "unknown variant `" + name + "` for geometry_type");
            }
        }
        sc.geometry = g;
        HostGeometry geo(g);
        if (opts.width <= 0 || opts.height <= 0) bad(GRRT_ERR_INVALID_CONFIGURATION, "--width and --height must be positive");

        // camera chart, four-velocity, future-directedness (cli/<geometry>.rs, cli/shared.rs:23-86)
        Vec4 pos = geo.camera_position_native(opts.camera_position);
        {
            hs.camera_velocity_mode = "StaticObserver";
            const TomlValue* body = nullptr;
            if (const TomlValue* cv = cfg.find("camera_velocity")) enum_variant(*cv, "camera_velocity", hs.camera_velocity_mode, body);
            if (hs.camera_velocity_mode == "Explicit") {
                const TomlValue& comps = need(*body, "components", "Explicit");
                if (comps.kind != TomlValue::ARRAY || comps.arr.size() != 4) bad(GRRT_ERR_TOML, "Explicit.components needs 4 numbers");
                for (int i = 0; i < 4; ++i) hs.camera_velocity_explicit[i] = comps.arr[i].as_double();
            }
        }
        Vec4 vel = resolve_camera_velocity(geo, pos, hs.camera_velocity_mode, hs.camera_velocity_explicit);

        // create_scene (cli/shared.rs:131-321)
        const TomlValue* adc = cfg.find("adaptive_sampling");
        adaptive(adc);
        sc.integration.max_steps = opts.max_steps;
        sc.integration.max_radius = opts.max_radius;
        sc.integration.step_size = opts.step_size;
        sc.integration.epsilon = opts.epsilon;
        sc.kerr_rhs_mode = GRRT_KERR_RHS_REFERENCE_FD;
        sc.celestial = texture(need(cfg, "celestial_texture", "RenderConfig"));

        make_camera(geo, pos, vel, opts, sc.camera);

        // objects
        if (const TomlValue* objs = cfg.find("objects")) {
            if (objs->kind != TomlValue::ARRAY) bad(GRRT_ERR_TOML, "objects must be an array of tables");
            for (const TomlValue& entry : objs->arr) {
                std::string kind;
                const TomlValue* b;
                enum_variant(entry, "objects", kind, b);
                if (sc.n_objects >= GRRT_MAX_OBJECTS) bad(GRRT_ERR_INVALID_CONFIGURATION, "too many objects (max 8)");
                GrrtObject& o = sc.objects[sc.n_objects];
                std::memset(&o, 0, sizeof o);
                if (kind == "Sphere") {
                    o.kind = GRRT_OBJ_SPHERE;
                    o.radius = need_f64(*b, "radius", "Sphere");
                    need_vec3(*b, "position", "Sphere", o.position);
                    o.texture = texture(need(*b, "texture", "Sphere"));
                    o.sphere_temperature = need_f64(*b, "temperature", "Sphere");
                } else if (kind == "Disc") {
                    o.kind = GRRT_OBJ_DISC;
                    o.inner_radius = need_f64(*b, "inner_radius", "Disc");
                    o.outer_radius = need_f64(*b, "outer_radius", "Disc");
                    o.texture = texture(need(*b, "texture", "Disc"));
                    o.temperature = temperature(g, need_f64(*b, "temperature", "Disc"), o.outer_radius);
                } else if (kind == "VolumetricDisc") {
                    const char* W = "VolumetricDisc";
                    o.kind = GRRT_OBJ_VOLUMETRIC_DISC;
                    o.inner_radius = need_f64(*b, "inner_radius", W);
                    o.outer_radius = need_f64(*b, "outer_radius", W);
                    double temp = need_f64(*b, "temperature", W);
                    o.num_octaves = need_usize(*b, "num_octaves", W);
                    o.vol_max_steps = need_usize(*b, "max_steps", W);
                    o.vol_step_size = need_f64(*b, "step_size", W);
                    o.thickness = need_f64(*b, "thickness", W);
                    o.density_multiplier = need_f64(*b, "density_multiplier", W);
                    o.brightness_reference_temperature = need_f64(*b, "brightness_reference_temperature", W);
                    o.absorption = need_f64(*b, "absorption", W);
                    o.scattering = need_f64(*b, "scattering", W);
                    need_vec3(*b, "noise_scale", W, o.noise_scale);
                    o.noise_offset = need_f64(*b, "noise_offset", W);
                    // validation, cli/shared.rs:229-275
                    auto req = [&](bool ok, const char* what) {
                        if (!ok) bad(GRRT_ERR_INVALID_CONFIGURATION, std::string("VolumetricDisc requires ") + what);
                    };
                    req(o.outer_radius > o.inner_radius, "outer_radius > inner_radius");
                    req(o.thickness > 0.0, "thickness > 0");
                    req(o.vol_max_steps != 0, "max_steps > 0");
                    req(o.vol_step_size > 0.0, "step_size > 0");
                    req(o.brightness_reference_temperature > 0.0, "brightness_reference_temperature > 0");
                    req(o.absorption >= 0.0, "absorption >= 0");
                    req(o.scattering >= 0.0, "scattering >= 0");
                    o.texture = texture(need(*b, "texture", W));
                    o.temperature = temperature(g, temp, o.outer_radius);
                    double axis[3] = {0.0, 0.0, 1.0};
                    if (b->find("axis")) need_vec3(*b, "axis", W, axis);
                    {  // VolumetricDisc::new (volumetric_disc.rs:61-73)
                        double n2 = (axis[0] * axis[0] + axis[1] * axis[1]) + axis[2] * axis[2];
                        double ax[3] = {0.0, 0.0, 1.0};
                        if (!(n2 <= 2.220446049250313e-16)) {
                            double n = std::sqrt(n2);
                            for (int i = 0; i < 3; ++i) ax[i] = axis[i] / n;
                        }
                        double base[3] = {1.0, 0.0, 0.0};
                        if (std::fabs(ax[0]) > 0.9) { base[0] = 0.0; base[1] = 1.0; }
                        auto cross = [](const double a[3], const double bb[3], double out[3]) {
                            out[0] = a[1] * bb[2] - a[2] * bb[1];
                            out[1] = a[2] * bb[0] - a[0] * bb[2];
                            out[2] = a[0] * bb[1] - a[1] * bb[0];
                        };
                        auto normalize = [](double v[3]) {
                            double n = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
                            for (int i = 0; i < 3; ++i) v[i] = v[i] / n;
                        };
                        double e1[3], e2[3];
                        cross(base, ax, e1);
                        normalize(e1);
                        cross(ax, e1, e2);
                        normalize(e2);
                        for (int i = 0; i < 3; ++i) {
                            o.axis[i] = ax[i];
                            o.e1[i] = e1[i];
                            o.e2[i] = e2[i];
                        }
                    }
                    uint32_t seed = 1;  // cli/shared.rs:294
                    if (const TomlValue* sv = b->find("perlin_seed")) {
                        if (sv->kind != TomlValue::INT || sv->i < 0 || sv->i > 4294967295ll) bad(GRRT_ERR_TOML, "perlin_seed must be a u32");
                        seed = (uint32_t)sv->i;
                    }
                    perlin_table(seed, o.perlin_perm);
                } else {
                    bad(GRRT_ERR_TOML, "unknown variant `" + kind + "` for objects");
                }
                ++sc.n_objects;
            }
        } else {
            bad(GRRT_ERR_TOML, "missing field `objects` in RenderConfig");
        }
        sc.celestial_temperature = need_f64(cfg, "celestial_temperature", "RenderConfig");
        finish();
    }

    // resolve indices into stable pointers (vectors no longer grow)
    void finish() {
        GrrtScene& sc = hs.scene;
        auto fix_tex = [&](GrrtTexture& t) {
            if (t.kind == GRRT_TEX_BITMAP) {
                t.rgba = hs.bitmaps[t._pad].rgba.data();
                t._pad = 0;
            }
        };
        fix_tex(sc.celestial);
        for (uint32_t k = 0; k < sc.n_objects; ++k) {
            GrrtObject& o = sc.objects[k];
            fix_tex(o.texture);
            if (o.kind != GRRT_OBJ_SPHERE && o.temperature.kind == GRRT_TEMP_LUT) {
                size_t idx = (size_t)o.temperature.constant;
                auto it = hs.temp_luts.begin();
                std::advance(it, idx);
                o.temperature.constant = 0.0;
                o.temperature.lut_r = it->r.data();
                o.temperature.lut_t = it->t.data();
            }
        }
        if (!hs.bb_log_t.empty()) {
            sc.bb_lut_n = GRRT_BLACKBODY_LUT_N;
            sc.bb_lut_log_t = hs.bb_log_t.data();
            sc.bb_lut_xyz = hs.bb_xyz.data();
        }
    }
};

int build_scene_from_toml(const std::string& text, const char* resource_root, const GrrtCliOpts& opts, HostSceneData& out,
                          std::string& err) {
    try {
        TomlValue cfg = parse_toml(text);
        SceneBuilder b(out, opts, resource_root);
        b.build(cfg);
        return GRRT_OK;
    } catch (const TomlError& e) {
        err = e.what();
        return GRRT_ERR_TOML;
    } catch (const ConfigError& e) {
        err = e.msg;
        return e.code;
    } catch (const std::exception& e) {
        err = e.what();
        return GRRT_ERR_INVALID_ARGUMENT;
    }
}

// A new camera for an assembled scene (animation frames): position, angles and frame size from `opts`, everything
// else -- textures, look-up tables, objects, integration limits -- stays as built.
int update_camera(HostSceneData& hsd, const GrrtCliOpts& opts, std::string& err) {
    try {
        if (opts.width <= 0 || opts.height <= 0) bad(GRRT_ERR_INVALID_CONFIGURATION, "--width and --height must be positive");
        HostGeometry geo(hsd.scene.geometry);
        Vec4 pos = geo.camera_position_native(opts.camera_position);
        Vec4 vel = resolve_camera_velocity(geo, pos, hsd.camera_velocity_mode, hsd.camera_velocity_explicit);
        GrrtCamera cam;
        std::memset(&cam, 0, sizeof cam);
        make_camera(geo, pos, vel, opts, cam);
        hsd.scene.camera = cam;
        return GRRT_OK;
    } catch (const ConfigError& e) {
        err = e.msg;
        return e.code;
    } catch (const std::exception& e) {
        err = e.what();
        return GRRT_ERR_INVALID_ARGUMENT;
    }
}

// render-ray-at (cli/schwarzschild.rs:89-121, cli/kerr.rs:77-100, cli/kerr_bl.rs:78-107, cli/euclidean*.rs)
int ray_at(const GrrtGeometry& g, const double position_cartesian[3], const double direction[3], double pos_out[4],
           double mom_out[4], std::string& err) {
    HostGeometry geo(g);
    Vec4 pos = geo.camera_position_native(position_cartesian);
    Frame T = geo.tetrad_at(pos);
    Vec4 mom;
    double d1 = direction[0], d2 = direction[1], d3 = direction[2];
    if (g.kind == GRRT_GEOM_EUCLIDEAN || g.kind == GRRT_GEOM_EUCLIDEAN_SPHERICAL) {
        // cli/euclidean.rs:74-99, cli/euclidean_spherical.rs:77-142: p = (|d|, d) in the chart's components
        double spatial_norm = std::sqrt(d1 * d1 + d2 * d2 + d3 * d3);
        if (!(std::isfinite(spatial_norm) && spatial_norm > 0.0)) {
            err = "render_ray_at direction must have a non-zero finite spatial part.";
            return GRRT_ERR_INVALID_CONFIGURATION;
        }
        if (g.kind == GRRT_GEOM_EUCLIDEAN) {
            mom = {spatial_norm, d1, d2, d3};
        } else {
            double r = pos[1], theta = pos[2], phi = pos[3];
            double sin_theta = std::sin(theta);
            if (!(std::isfinite(r) && r > 0.0) || !(std::isfinite(sin_theta) && std::fabs(sin_theta) > 1e-12)) {
                err = "Euclidean-spherical render_ray_at requires r > 0 away from the polar axis";
                return GRRT_ERR_INVALID_CONFIGURATION;
            }
            double r_dot = sin_theta * std::cos(phi) * d1 + sin_theta * std::sin(phi) * d2 + std::cos(theta) * d3;
            double theta_unit_dot = std::cos(theta) * std::cos(phi) * d1 + std::cos(theta) * std::sin(phi) * d2 - sin_theta * d3;
            double phi_unit_dot = -std::sin(phi) * d1 + std::cos(phi) * d2;
            mom = {spatial_norm, r_dot, theta_unit_dot / r, phi_unit_dot / (r * sin_theta)};
        }
    } else if (g.kind == GRRT_GEOM_SCHWARZSCHILD) {
        double theta = pos[2], phi = pos[3];
        double r_d = std::sin(theta) * std::cos(phi) * d1 + std::sin(theta) * std::sin(phi) * d2 + std::cos(theta) * d3;
        double theta_d = std::cos(theta) * std::cos(phi) * d1 + std::cos(theta) * std::sin(phi) * d2 - std::sin(theta) * d3;
        double phi_d = -std::sin(phi) * d1 + std::cos(phi) * d2;
        mom = add(add(add(scale(1.0, T.t), scale(phi_d, T.x)), scale(-theta_d, T.y)), scale(-r_d, T.z));
    } else {
        Vec4 space = add(add(scale(d1, T.x), scale(d2, T.y)), scale(d3, T.z));
        double n = std::sqrt(geo.inner(pos, space, space));  // Kerr / KerrBL: signature (-+++)
        auto term = [&](const Vec4& e, double d) {
            Vec4 s = scale(d, e);
            return Vec4{s[0] / n, s[1] / n, s[2] / n, s[3] / n};
        };
        mom = add(add(add(scale(1.0, T.t), term(T.x, d1)), term(T.y, d2)), term(T.z, d3));
    }
    double orientation = geo.sig0() * geo.inner(pos, T.t, mom);
    if (!(std::isfinite(mom[0]) && orientation > 0.0)) {
        err = "render_ray_at momentum has wrong time orientation";
        return GRRT_ERR_INVALID_CONFIGURATION;
    }
    for (int i = 0; i < 4; ++i) {
        pos_out[i] = pos[i];
        mom_out[i] = mom[i];
    }
    return GRRT_OK;
}

// Point::to_cartesian (point.rs:140-155) for the CSV of ray.rs:35-54
void step_to_cartesian(const GrrtGeometry& g, const double x[4], double out[4]) {
    HostGeometry geo(g);
    out[0] = x[0];
    switch (geo.chart()) {
        case CART:
            out[1] = x[1]; out[2] = x[2]; out[3] = x[3];
            break;
        case SPH:
            out[1] = x[1] * std::sin(x[2]) * std::cos(x[3]);
            out[2] = x[1] * std::sin(x[2]) * std::sin(x[3]);
            out[3] = x[1] * std::cos(x[2]);
            break;
        default:
            out[1] = (x[1] * std::cos(x[3]) - g.a * std::sin(x[3])) * std::sin(x[2]);
            out[2] = (x[1] * std::sin(x[3]) + g.a * std::cos(x[3])) * std::sin(x[2]);
            out[3] = x[1] * std::cos(x[2]);
            break;
    }
}

}  // namespace grrt_host
