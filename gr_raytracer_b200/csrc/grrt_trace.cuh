// grrt_trace.cuh -- the persistent ray-tracing kernel (K1/K3 of SURVEY.md section 2).
//
// Scheduling: a fixed grid of persistent CTAs; every lane owns one ray at a time and pulls the
// next work unit from a global atomic counter when its ray has terminated (one atomicAdd per
// warp, lanes ranked with ballot/popc).  All lanes of a warp therefore execute the same RKF45
// attempt code regardless of which ray or step they are on, so the 100x spread in steps-per-ray
// (171 ... 15 500 ... 276 733) never idles a lane for long.
//
// Hot loop / events: the kernel is two nested loops.  The inner one runs RKF45 attempts until some
// lane of the warp has an event (an accepted step whose chord may touch an object or the celestial
// sphere, a ray that ended, an idle lane); everything that handles events -- Cartesian end points and
// the chord tests of Objects::intersects, the escape test, shading, the ray-march, terminal colour,
// write-out, ray set-up -- is outside it.  Inside the loop an accepted step only runs the
// conservative pre-filter segment_needs_cartesian() and the cheap stop tests.  Objects::intersects
// is split accordingly: its geometric half records which objects became the nearest hit at their
// turn; its shading half runs on them, in the same order, so that emitter errors fail the pixel
// exactly when the reference's `?` does.
//
// RKF45 attempt: the stage derivatives k1..k5 live in shared memory ([stage][component][thread],
// conflict-free).  State components whose derivative is identically zero (KerrBL: 6,7;
// Kerr-Schild: p_t; flat Cartesian: all momenta) are skipped: they contribute exact zeros to every
// sum of the reference, so nothing changes numerically; and only the components the right-hand
// side reads have their k stored.  The exact build (-fmad=false; mandatory for the 2000-instruction
// Kerr-Schild right-hand side) runs the six stages as a real loop -- one copy of the right-hand side
// in the instruction stream -- and accumulates the 5th-order solution and the error vector as each
// stage completes, in the order runge_kutta.rs:115-123 sums them.  The FMA build runs the stages as
// straight-line code (GRRT_UNROLL_STAGES), takes sin / cos of a stage's polar angle from the state's
// by the addition theorem (GRRT_SINCOS_DELTA) and forms y_new / the error vector of the stored
// components at the last stage from the k1..k5 it loads anyway (GRRT_DEFER_SUMS).  The unrolled
// KerrBL loop is instruction-fetch bound: rare paths inside it are calls.
//
// The reference stores the whole trajectory (112 B per step) and walks it a second time
// (scene.rs:141-152); here each accepted step is tested against the objects as it is produced
// and only the previous state is kept (in shared memory: it is read only when a hit is shaded).
//
// VolumetricDisc ray-marches are executed warp-cooperatively: the lane that found the hit hands
// the march over through shared memory, the 32 lanes evaluate 32 consecutive samples in parallel
// (density, Perlin fBm, redshift, LUT colour) and the order-dependent accumulation (transparency
// product, emission sums) is then replayed in sample order so the result is the same sequence of
// floating-point operations the reference performs (volumetric_disc.rs:209-328).
//
// Multi-GPU: with TraceArgs::n_peers > 0 a finished pixel is stored into every rank's frame buffer
// (peer-mapped over NVLink) instead of a local compact buffer, so no gather pass follows.
#pragma once

#include "grrt_device.cuh"
#include "grrt_pow.cuh"

namespace GRRT_NS {

constexpr int kMaxHits = 32;
constexpr int kMaxRetryStep = 100;  // runge_kutta.rs:60
constexpr unsigned kFull = 0xffffffffu;

// Fehlberg weights as data for the stage loop (same expressions as the RK_* constants; the B coefficients are
// compile-time operands, see rk_b below)
__constant__ double c_rk_ch[6] = {RK_CH1, RK_CH2, RK_CH3, RK_CH4, RK_CH5, RK_CH6};
__constant__ double c_rk_ct[6] = {RK_CT1, RK_CT2, RK_CT3, RK_CT4, RK_CT5, RK_CT6};

// the B coefficients as data as well: FP64 instructions take 64-bit constants only from the constant bank (an immediate
// costs two UMOV), so the unrolled stage code of the FMA build reads them as c[bank][offset] operands
__constant__ double c_rk_b[25] = {RK_B21, 0, 0, 0, 0, RK_B31, RK_B32, 0, 0, 0, RK_B41, RK_B42, RK_B43, 0, 0,
                                  RK_B51, RK_B52, RK_B53, RK_B54, 0, RK_B61, RK_B62, RK_B63, RK_B64, RK_B65};

// which of the 8 state components have a non-zero derivative
template <int G>
struct Comp {
    __host__ __device__ static constexpr bool on(int i) {
        return G == GRRT_GEOM_EUCLIDEAN ? i < 4 : (G == GRRT_GEOM_KERR_BL ? i < 6 : (G == GRRT_GEOM_KERR ? i != 4 : true));
    }
    // ... and which the right-hand side reads: only those need their stage derivatives k1..k5 kept, because the
    // stage argument y + sum B k is dead for every other component (t and phi in the spherical charts, the
    // positions in flat Cartesian space)
    __host__ __device__ static constexpr bool read(int i) {
        return G == GRRT_GEOM_EUCLIDEAN ? i >= 4
             : G == GRRT_GEOM_KERR_BL  ? (i == 1 || i == 2 || i == 4 || i == 5)
             : G == GRRT_GEOM_KERR     ? i != 0
                                       : (i != 0 && i != 3);
    }
    __host__ __device__ static constexpr bool stored(int i) { return on(i) && read(i); }
    __host__ __device__ static constexpr int slot(int i) {  // dense index of a stored component
        int n = 0;
        for (int j = 0; j < i; ++j) n += stored(j) ? 1 : 0;
        return n;
    }
    static constexpr int count = slot(8);
};

// Fehlberg B coefficients as compile-time values (stage s = 1..5, term j < s)
__host__ __device__ constexpr double rk_b(int s, int j) {
    return s == 1 ? RK_B21
         : s == 2 ? (j == 0 ? RK_B31 : RK_B32)
         : s == 3 ? (j == 0 ? RK_B41 : j == 1 ? RK_B42 : RK_B43)
         : s == 4 ? (j == 0 ? RK_B51 : j == 1 ? RK_B52 : j == 2 ? RK_B53 : RK_B54)
                  : (j == 0 ? RK_B61 : j == 1 ? RK_B62 : j == 2 ? RK_B63 : j == 3 ? RK_B64 : RK_B65);
}
__host__ __device__ constexpr double rk_ch(int s) {
    return s == 0 ? RK_CH1 : s == 1 ? RK_CH2 : s == 2 ? RK_CH3 : s == 3 ? RK_CH4 : s == 4 ? RK_CH5 : RK_CH6;
}
__host__ __device__ constexpr double rk_ct(int s) {
    return s == 0 ? RK_CT1 : s == 1 ? RK_CT2 : s == 2 ? RK_CT3 : s == 3 ? RK_CT4 : s == 4 ? RK_CT5 : RK_CT6;
}
// argument of stage SS: y + B_s1 k1 + B_s2 k2 + ... summed left to right (runge_kutta.rs:95-124); straight-line so
// that all shared-memory loads of a stage are issued together and the coefficients are instruction operands
#if defined(GRRT_FAST_MATH) && (defined(GRRT_UNROLL_STAGES) || defined(GRRT_RKB_CONST))
#define RKB(SS, j) c_rk_b[((SS) - 1) * 5 + (j)]
#else
#define RKB(SS, j) rk_b(SS, j)
#endif
template <int G, int SS>
__device__ __forceinline__ void stage_argument(const double* kcol, const double* y, double* tmp) {
    constexpr int NC = Comp<G>::count;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double acc = y[i];
        if (Comp<G>::stored(i)) {
#if defined(GRRT_SINCOS_DELTA) && defined(GRRT_FAST_MATH)
            if (i == 2 && GeomTraits<G>::chart != CH_CART) {
                // polar angle: the increment sum B k alone (the stage takes sin, cos of theta from it by the addition
                // theorem, see the stage loop); tmp[2] is not read by the right-hand side in these charts
                acc = RKB(SS, 0) * kcol[(0 * NC + Comp<G>::slot(i)) * kTraceThreads];
#pragma unroll
                for (int j = 1; j < SS; ++j) acc = acc + RKB(SS, j) * kcol[(j * NC + Comp<G>::slot(i)) * kTraceThreads];
                tmp[i] = acc;
                continue;
            }
#endif
#pragma unroll
            for (int j = 0; j < SS; ++j) acc = acc + RKB(SS, j) * kcol[(j * NC + Comp<G>::slot(i)) * kTraceThreads];
        }
        tmp[i] = acc;
    }
}

#ifdef GRRT_FAST_MATH
#define RKCH(j) c_rk_ch[j]
#define RKCT(j) c_rk_ct[j]
#else
#define RKCH(j) rk_ch(j)
#define RKCT(j) rk_ct(j)
#endif
// Argument of the last stage AND, from the same five shared-memory loads per component, the sums y + CH1 k1 + ... + CH5 k5
// and CT1 k1 + ... + CT5 k5 of the stored components (left to right, as runge_kutta.rs:115-123 forms them once all six k
// exist): the 5th-order solution and the error vector of those components then need no registers while stages 1..5
// run -- they are completed with k6 after the last right-hand side.
template <int G>
__device__ __forceinline__ void stage_argument_final(const double* kcol, const double* y, double* tmp, double* y_new, double* acc_e) {
    constexpr int NC = Comp<G>::count;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (!Comp<G>::stored(i)) {
            tmp[i] = y[i];
            continue;
        }
        double k[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) k[j] = kcol[(j * NC + Comp<G>::slot(i)) * kTraceThreads];
        double acc;
#if defined(GRRT_SINCOS_DELTA) && defined(GRRT_FAST_MATH)
        if (i == 2 && GeomTraits<G>::chart != CH_CART) {
            acc = RKB(5, 0) * k[0];
#pragma unroll
            for (int j = 1; j < 5; ++j) acc = acc + RKB(5, j) * k[j];
        } else
#endif
        {
            acc = y[i];
#pragma unroll
            for (int j = 0; j < 5; ++j) acc = acc + RKB(5, j) * k[j];
        }
        tmp[i] = acc;
        double yn = y[i], ae = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
#ifdef GRRT_FAST_MATH
            if (j == 1) continue;   // CH2 = CT2 = 0
#endif
            yn = yn + RKCH(j) * k[j];
            ae = ae + RKCT(j) * k[j];
        }
        y_new[i] = yn;
        acc_e[i] = ae;
    }
}

struct RayCtx {  // per-lane ray bookkeeping kept in registers
    unsigned long long unit;
    unsigned step;
    unsigned attempts;
    int n_hits;
    int retries;
    int prev_valid;   // the shared-memory copy of the previous step's Cartesian point is current
};
// ... and the part that is only touched at accepted steps / hits / ray end, kept in shared memory
// ([slot][thread], conflict-free) so that the RKF45 stage loop has the registers to itself
enum { CX_FREQ_E = 0, CX_FREQ_PT = 1, CX_FREQ_PPHI = 2, CX_OPACITY = 3, CX_PREV_X = 4, CX_PREV_Y = 5, CX_PREV_Z = 6, CX_SLOTS = 7,
       // volumetric kernels only: hit point and chord direction handed to the warp-cooperative ray-march
       CX_VOL_PX = 7, CX_VOL_PY = 8, CX_VOL_PZ = 9, CX_VOL_DX = 10, CX_VOL_DY = 11, CX_VOL_DZ = 12, CX_SLOTS_VOL = 13 };
struct Freq {
    double e, pt, pphi;
};

template <int G>
GD void camera_ray(const DScene& S, double row, double column, double pos[4], double mom[4]) {  // camera.rs:214-254
    double shifted_column = column + 1.0;
    double shifted_row = row + 1.0;
    double i_prime = S.cam_hand * S.cam_scale * (shifted_column - (S.cam_cols + 1.0) / 2.0);
    double j_prime = S.cam_scale * (shifted_row - (S.cam_rows + 1.0) / 2.0);
    double w_squared = S.cam_sig * (1.0 + i_prime * i_prime + j_prime * j_prime);
    double den = S.cam_sig * w_squared;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double w = (S.tet[3][i] + i_prime * S.tet[1][i]) + j_prime * S.tet[2][i];
        double dir = -S.tet[3][i] + (2.0 * w) / den;
        mom[i] = dir + (-S.tet[0][i]);
        pos[i] = S.cam_pos[i];
    }
}

// get_geodesic_solver + create_initial_state + get_ray_frequency_data
template <int G>
GD void init_ray(const DScene& S, const double pos[4], const double mom[4], double y[8], RayConst& C, Freq& F) {
    C.e = C.l_z = C.q = 0.0;
    if (G == GRRT_GEOM_KERR_BL) {  // kerr_bl.rs:505-577 (BL branch), 176-223
        double r = pos[1];
        double st, ct;
        m_sincos(pos[2], &st, &ct);
        BlMetric g = bl_metric(S.radius, S.a, r, st, ct);
        // p_cov = g * p with nalgebra's accumulation (zero entries kept)
        double pc0 = ((g.tt * mom[0] + 0.0 * mom[1]) + 0.0 * mom[2]) + g.tph * mom[3];
        double pc2 = ((0.0 * mom[0] + 0.0 * mom[1]) + g.thth * mom[2]) + 0.0 * mom[3];
        double pc3 = ((g.tph * mom[0] + 0.0 * mom[1]) + 0.0 * mom[2]) + g.phph * mom[3];
        double e = -pc0, l_z = pc3, p_theta = pc2;
        double sin2 = st * st;
        double q = p_theta * p_theta + ct * ct * (l_z * l_z / fmax(sin2, 1e-28) - S.a * S.a * e * e);
        C.e = e; C.l_z = l_z; C.q = q;
        double a = S.a;
        double sign_r = mom[1] >= 0.0 ? 1.0 : -1.0;
        double sign_theta = mom[2] >= 0.0 ? 1.0 : -1.0;
        double del = bl_delta(r, S.radius, a);
        double p_r = (r * r + a * a) * e - a * l_z;
        double lae = l_z - a * e;
        double r_pot = p_r * p_r - del * (lae * lae + q);
        double th_pot = q + a * a * e * e * ct * ct - l_z * l_z * ct * ct / (st * st);
        y[0] = pos[0]; y[1] = r; y[2] = pos[2]; y[3] = pos[3];
        y[4] = sign_r * sqrt(fmax(r_pot, 0.0));
        y[5] = sign_theta * sqrt(fmax(th_pot, 0.0));
        y[6] = 0.0; y[7] = 0.0;
    } else if (G == GRRT_GEOM_KERR) {  // kerr.rs:243-260
        double g10[10], m[4][4], pc[4];
        ks_metric10(S.radius, S.a, pos[1], pos[2], pos[3], g10);
        sym10_to_16(g10, m);
        mat4_vec(m, mom, pc);
        y[0] = pos[0]; y[1] = pos[1]; y[2] = pos[2]; y[3] = pos[3];
        y[4] = pc[0]; y[5] = pc[1]; y[6] = pc[2]; y[7] = pc[3];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            y[i] = pos[i];
            y[4 + i] = mom[i];
        }
    }
    // redshift.rs:45-60
    double e_t[4] = {1.0, 0.0, 0.0, 0.0};
    double axial[4] = {0.0, 0.0, 0.0, 1.0};
    if (GeomTraits<G>::chart == CH_CART) {
        axial[1] = -pos[2]; axial[2] = pos[1]; axial[3] = 0.0;
    }
    F.e = inner_product<G>(S, pos, S.cam_vel, mom);
    F.pt = inner_product<G>(S, pos, e_t, mom);
    F.pphi = inner_product<G>(S, pos, axial, mom);
}

// does_exit of a march without a cached exit distance (volumetric_disc.rs:154-170): four cylinder / cap intersections
// per sample, out of line so that the march loop of the common (cached) case stays compact
__device__ __noinline__ bool vol_does_exit_cold(const DObject& o, V3 p, V3 rd, double d_s) {
    V3 to = v3(p.x + rd.x * d_s, p.y + rd.y * d_s, p.z + rd.z * d_s);
    double t;
    return vol_first_hit(o, p, to, t) && in01(t);
}

// ---- warp-cooperative VolumetricDisc ray-march (volumetric_disc.rs:209-328) --------------------------
// Every argument is warp-uniform.  Returns the march colour; *err receives a GRRT_RAYERR_* when
// compute_temperature failed (the caller turns that into (0,0,0,0) as color_at_uv does, :596-600).
template <int G>
__device__ __forceinline__ Cie vol_march_warp(const DScene& S, const DObject& o, const uint8_t* perm, V3 ro, V3 rd,
                                           double e_obs, double p_t, double p_phi, int* err,
                                           unsigned long long* samples) {
    const int lane = threadIdx.x & 31;
    const double sigma = o.absorption + o.scattering;
    const double d_s = o.vol_step;
    const V3 axis = obj_axis(o);
    *err = 0;
    // Empty marches.  Every chord that LEAVES the capture surface also triggers a march (volumetric_disc.rs:530-549
    // takes the first hit with t > 1e-9); it starts on the surface, points away from the gas, finds no cached exit
    // and runs all max_steps samples through zero density -- ~98 % of all samples of configs 4 and 5.  Its result
    // is exactly (0,0,0,0): no sample attenuates or emits, so transparency stays 1 and alpha = (1 - 1) * 1.
    // Whether ANY sample can see gas is decided here, conservatively, from the geometry of the density support
    // (compute_density, :97-138: r in (r_in, r_out) and exp(-(h/thickness)^2) >= 0.001, i.e. |h| < 2.6283 thickness):
    // if the ray segment [0, max_distance] stays outside the slab |h| < 2.7 thickness, or within the slab stays
    // inside the hole / beyond the rim, the march is skipped with the result the reference would have produced.
    {
        const double smax = o.vol_step * (double)o.vol_max_steps * (1.0 + 1e-9) + o.vol_step;
        const double Hd = 2.7 * o.thickness;
        const double h0 = dot3(ro, axis), hd = dot3(rd, axis);
        double sa = 0.0, sb = smax;  // sub-interval of the ray inside the slab
        bool empty = false;          // set only by comparisons that are false for NaN: a doubtful ray is marched
        if (fabs(hd) < 1e-300) {
            empty = fabs(h0) >= Hd;
        } else {
            double s1 = (-Hd - h0) / hd, s2 = (Hd - h0) / hd;
            if (s1 > s2) { double t = s1; s1 = s2; s2 = t; }
            if (s1 - 1e-9 * (1.0 + fabs(s1)) > sa) sa = s1 - 1e-9 * (1.0 + fabs(s1));
            if (s2 + 1e-9 * (1.0 + fabs(s2)) < sb) sb = s2 + 1e-9 * (1.0 + fabs(s2));
            empty = sa > sb;
        }
        if (!empty) {  // rho^2(s) = |(ro + s rd) x axis|^2 = A s^2 + B s + C over [sa, sb]
            const V3 v = cross3(ro, axis), w = cross3(rd, axis);
            const double A = dot3(w, w), B = 2.0 * dot3(v, w), C0 = dot3(v, v);
            const double qa = (A * sa + B) * sa + C0, qb = (A * sb + B) * sb + C0;
            double qmin = fmin(qa, qb);
            const double qmax = fmax(qa, qb);  // convex in s: the maximum is at an end point
            if (A > 0.0) {
                const double sv = -B / (2.0 * A);
                if (sv > sa && sv < sb) qmin = fmin(qmin, (A * sv + B) * sv + C0);
            }
            const double scale = 1e-9 * (fabs(C0) + fabs(B) * smax + A * smax * smax + 1.0);
            empty = (qmax + scale < o.r_in * o.r_in) || (qmin - scale > o.r_out * o.r_out);
        }
        if (empty) return cie(0.0, 0.0, 0.0, 0.0);
    }
    // precompute_exit_distance (:172-196)
    double exit_distance = 0.0;
    bool cached;
    {
        double max_distance = o.vol_step * (double)o.vol_max_steps;
        V3 to = ro + rd * max_distance;
        double t;
        cached = vol_first_hit(o, ro, to, t);
        if (cached) exit_distance = t * max_distance;
    }
    double transparency = 1.0, aws = 0.0, awt = 0.0, ax = 0.0, ay = 0.0, az = 0.0;
    double D = 0.0;  // d_o at the first sample of the current block of 32
    const unsigned long long max_steps = o.vol_max_steps;
    *err = 0;
    for (unsigned long long base = 0; base < max_steps; base += 32) {
        double d_o = D;  // d_o is accumulated by repeated addition in the reference: replay it
        for (int j = 0; j < 31; ++j)
            if (j < lane) d_o += d_s;
        const bool valid = base + (unsigned long long)lane < max_steps;
        V3 p = ro + rd * d_o;
        double d_o_next = d_o + d_s;
        double density = valid ? vol_density(o, perm, p) : 0.0;
        const bool has = density > 0.0;
        double att = 1.0, light_att = 0.0, intensity = 0.0;
        Cie light = cie(0.0, 0.0, 0.0, 0.0);
        bool emit = false;
        int terr = 0;
        if (has) {
            att = m_exp(-d_s * density * sigma);
            double u_t, u_phi;
            if (orbit_coefficients_cart<G>(S, p, u_t, u_phi)) {
                double emitter_energy = u_t * p_t + u_phi * p_phi;
                double redshift = e_obs / emitter_energy;
                double r_dist = norm3(cross3(p, axis));
                double temperature;
                terr = compute_temperature(o.temp, r_dist, temperature);
                if (!terr) {
                    double u, v;
                    vol_get_uv(o, p, u, v);
                    light = texture_color(S, o.tex, u, v, temperature, redshift);
                    light_att = m_exp(-density * d_s * sigma);
                    double tr = temperature / o.bref_temperature;
                    double tr2 = tr * tr;
                    intensity = tr2 * tr2;
                    emit = true;
                }
            }
        }
        bool exited = false;
        if (valid) {
            if (cached) {
                exited = d_o_next >= exit_distance;
            } else {  // does_exit (:154-170)
                exited = vol_does_exit_cold(o, p, rd, d_s);
            }
        }
        const unsigned m_valid = __ballot_sync(kFull, valid);
        const unsigned m_exit = __ballot_sync(kFull, exited);
        const unsigned m_err = __ballot_sync(kFull, terr != 0);
        const unsigned m_has = __ballot_sync(kFull, has);
        const unsigned m_emit = __ballot_sync(kFull, emit);
        const int first_exit = m_exit ? (__ffs(m_exit) - 1) : 32;
        const int first_err = m_err ? (__ffs(m_err) - 1) : 32;
        const int last = min(first_exit, __popc(m_valid) - 1);  // last sample of this block that runs
        if (first_err <= last) {
            *err = __shfl_sync(kFull, terr, first_err);
            *samples += (unsigned long long)(first_err + 1);
            return cie(0.0, 0.0, 0.0, 0.0);
        }
        const unsigned upto = (last >= 31) ? kFull : ((1u << (last + 1)) - 1u);
        // transparency *= sample_attenuation, in sample order
        double my_t = transparency;
        for (unsigned mm = m_has & upto; mm; mm &= mm - 1) {
            int j = __ffs(mm) - 1;
            transparency *= __shfl_sync(kFull, att, j);
            if (j == lane) my_t = transparency;
        }
        double cx = 0.0, cy = 0.0, cz = 0.0, asw = 0.0, aw = 0.0;
        if (emit) {
            double emission_weight = my_t * light_att * o.scattering * density * d_s;
            double f = emission_weight * intensity;
            cx = light.x * f;
            cy = light.y * f;
            cz = light.z * f;
            aw = density * d_s;
            asw = rclamp(light.a, 0.0, 1.0) * aw;
        }
        for (unsigned mm = m_emit & upto; mm; mm &= mm - 1) {
            int j = __ffs(mm) - 1;
            aws += __shfl_sync(kFull, asw, j);
            awt += __shfl_sync(kFull, aw, j);
            ax += __shfl_sync(kFull, cx, j);
            ay += __shfl_sync(kFull, cy, j);
            az += __shfl_sync(kFull, cz, j);
        }
        *samples += (unsigned long long)(last + 1);
        if (first_exit <= last) break;
        D = __shfl_sync(kFull, d_o_next, 31);
    }
    double physical_opacity = 1.0 - transparency;
    double texture_alpha = awt > 0.0 ? aws / awt : 1.0;
    return cie(ax, ay, az, physical_opacity * texture_alpha);
}

// ---- the kernel ----------------------------------------------------------------------------------------
enum { ST_IDLE = 0, ST_ACTIVE = 1, ST_DEAD = 2 };

#ifdef GRRT_FAST_MATH
// The step-size controller's exact path is cold in the fast build (seed out of range, err = 0 / inf / NaN): kept out
// of line so that the attempt loop stays short.
__device__ __noinline__ double controller_pow_cold(double epsilon, double err) {
    return pow_glibc(pow_tables_global(), epsilon / err, 1.0 / 5.0);
}
#endif

// the exact build keeps the glibc-compatible pow tables in shared memory (one pow per attempt); the fast build only
// needs them on the controller's rare fallback path and reads them from global memory
#ifdef GRRT_FAST_MATH
constexpr size_t kPowSmemBytes = 0;
#else
constexpr size_t kPowSmemBytes = sizeof(double) * 128 * 3 + 8 * 256;
#endif
template <int G, bool HAS_VOL>
constexpr size_t trace_smem_bytes() {
    return sizeof(double) * (size_t)(5 * Comp<G>::count + 8 + (HAS_VOL ? CX_SLOTS_VOL : CX_SLOTS)) * kTraceThreads + kPowSmemBytes +
           (HAS_VOL ? GRRT_MAX_OBJECTS * 256 : 0);
}

template <int G, bool HAS_VOL, int MIN_BLOCKS>
__global__ void __launch_bounds__(kTraceThreads, MIN_BLOCKS)
trace_kernel(const __grid_constant__ DScene S, const __grid_constant__ TraceArgs A) {
    constexpr int NC = Comp<G>::count;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_k = reinterpret_cast<double*>(smem_raw);                       // [5][NC][threads]
    double* s_yprev = s_k + 5 * NC * kTraceThreads;                          // [8][threads]
    double* s_ctx = s_yprev + 8 * kTraceThreads;                             // [CX_SLOTS][threads]
    unsigned char* s_tail = reinterpret_cast<unsigned char*>(s_ctx + (HAS_VOL ? CX_SLOTS_VOL : CX_SLOTS) * kTraceThreads);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
#ifndef GRRT_FAST_MATH
    PowTables pow_tab;
    double (*s_pow_log)[3] = reinterpret_cast<double (*)[3]>(s_tail);        // glibc-compatible pow tables
    unsigned long long* s_pow_exp = reinterpret_cast<unsigned long long*>(s_pow_log + 128);
    for (int i = tid; i < 128 * 3; i += kTraceThreads) (&s_pow_log[0][0])[i] = (&kPowLogTab[0][0])[i];
    for (int i = tid; i < 256; i += kTraceThreads) s_pow_exp[i] = kExpTab[i];
    pow_tab.log_tab = s_pow_log;
    pow_tab.exp_tab = s_pow_exp;
#endif
    uint8_t* s_perm = reinterpret_cast<uint8_t*>(s_tail + kPowSmemBytes);    // [n_vol][256]
    if (HAS_VOL) {
        for (int k = 0; k < S.n_objects; ++k)
            if (S.obj[k].kind == GRRT_OBJ_VOLUMETRIC_DISC)
                for (int i = tid; i < 256; i += kTraceThreads) s_perm[S.obj[k].perm_slot * 256 + i] = S.obj[k].perm[i];
    }
    __syncthreads();

    int state = ST_IDLE;
    double y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = 0.0;
    double h = 0.0, h_cur = 0.0;
    RayConst C = {0.0, 0.0, 0.0};
    RayCtx R;
    R.n_hits = 0; R.retries = 0; R.unit = 0; R.step = 0; R.attempts = 0; R.prev_valid = 0;
#define CTX(slot) s_ctx[(slot) * kTraceThreads + tid]
    Cie hits[kMaxHits];
    const unsigned ss = A.spa ? A.spa * A.spa : 1u;

    for (;;) {
        // ---------------- refill idle lanes from the global queue ----------------
        const unsigned need = __ballot_sync(kFull, state == ST_IDLE);
        if (need) {
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(A.queue, (unsigned long long)__popc(need));
            base = __shfl_sync(kFull, base, leader);
            if (state == ST_IDLE) {
                const unsigned long long unit = base + (unsigned long long)__popc(need & ((1u << lane) - 1u));
                if (unit >= A.n_units) {
                    state = ST_DEAD;
                } else {
                    // unit -> (item, stratum) -> pixel
                    unsigned long long item = unit / ss;
                    unsigned k = (unsigned)(unit % ss);
                    unsigned long long pid;
                    bool ok = true;
                    if (A.tiled) {
                        unsigned long long tile = unit >> 5;
                        unsigned in_tile = (unsigned)(unit & 31);
                        unsigned px = (unsigned)(tile % A.tiles_x) * 8u + (in_tile & 7u);
                        unsigned py = (unsigned)(tile / A.tiles_x) * 4u + (in_tile >> 3);
                        ok = px < A.width && py < A.height;
                        pid = (unsigned long long)py * A.width + px;
                        item = pid;
                    } else {
                        pid = A.items ? (unsigned long long)A.items[A.item_offset + item * A.item_stride] : A.first_item + item;
                    }
                    if (ok) {
                        const long long row = (long long)(pid / A.width) + A.from_row;
                        const long long col = (long long)(pid % A.width) + A.from_col;
                        double frow = (double)row, fcol = (double)col;
                        if (A.spa) {  // stratified_sample_offset + get_ray_for_offset (raytracer.rs:145-159, camera.rs:247-254)
                            unsigned sr = k / A.spa, sc = k % A.spa;
                            unsigned long long sample_index = (unsigned long long)sr * A.spa + sc;
                            double dx = ((double)sc + hash_pixel_samples(row, col, 2ull * sample_index)) / (double)A.spa;
                            double dy = ((double)sr + hash_pixel_samples(row, col, 2ull * sample_index + 1ull)) / (double)A.spa;
                            frow = (double)row + (dy - 0.5);
                            fcol = (double)col + (dx - 0.5);
                        }
                        double pos[4], mom[4];
                        camera_ray<G>(S, frow, fcol, pos, mom);
                        Freq F;
                        init_ray<G>(S, pos, mom, y, C, F);
                        CTX(CX_FREQ_E) = F.e;
                        CTX(CX_FREQ_PT) = F.pt;
                        CTX(CX_FREQ_PPHI) = F.pphi;
                        CTX(CX_OPACITY) = 0.0;
                        const V3 c0 = position_cartesian<G>(S, y);
                        CTX(CX_PREV_X) = c0.x;
                        CTX(CX_PREV_Y) = c0.y;
                        CTX(CX_PREV_Z) = c0.z;
                        R.unit = (A.tiled || A.n_peers) ? pid : unit;  // index of the ray's result in the output buffers
                        R.step = 0;
                        R.attempts = 0;
                        R.n_hits = 0;
                        R.retries = 0;
                        R.prev_valid = 1;
                        h = S.step_size;
                        h_cur = rclamp(h, kHMin, kHMax);
                        state = ST_ACTIVE;
                    }
                }
            }
        }
        if (__ballot_sync(kFull, state != ST_DEAD) == 0) break;

        const bool active = state == ST_ACTIVE;
        int finish = -1;       // >= 0: GRRT_STOP_* with which this lane's ray ends
        int ray_error = 0;
        unsigned cand = 0;     // objects that were the nearest hit so far at their turn in Objects::intersects
        bool geo_pending = false;  // the step just accepted needs the chord / escape tests of the event code
        if (active && S.max_steps <= 1) finish = GRRT_STOP_NONE;  // `for i in 1..max_steps` runs zero times
        bool event = state == ST_IDLE || finish >= 0;

        // ================= hot loop: RKF45 attempts until some lane of the warp has an event =================
        // An event is a chord that hit an object (to be shaded), the end of a ray, or an idle lane.  Everything
        // that handles events -- shading, the ray-march, terminal colour, write-out, fetching and launching the next
        // ray -- sits outside this loop, so the loop body is one contiguous piece of code of ~1.5k instructions.
        while (!__any_sync(kFull, event)) {
            if (active) {
                // ---------------- one RKF45 attempt (runge_kutta.rs:86-182, flattened retry loop) ----------------
                double y_new[8];
                bool accepted = false;
                double acc_e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#ifdef GRRT_DEFER_SUMS
                    if (Comp<G>::stored(i)) continue;   // formed at the last stage from the stored k1..k5
#endif
                    y_new[i] = y[i];
                    acc_e[i] = 0.0;
                }
#if defined(GRRT_SINCOS_DELTA) && defined(GRRT_FAST_MATH)
                double sin0 = 0.0, cos0 = 1.0;
#endif
                double* const kcol = s_k + tid;  // this lane's column of the [stage][component][thread] k store
                double* kst = kcol;
#ifdef GRRT_UNROLL_STAGES
#pragma unroll
#else
#pragma unroll 1
#endif
                for (int s = 0; s < 6; ++s) {
                    double tmp[8];
                    switch (s) {
                        case 0:
#pragma unroll
                            for (int i = 0; i < 8; ++i) tmp[i] = y[i];
                            break;
                        case 1: stage_argument<G, 1>(kcol, y, tmp); break;
                        case 2: stage_argument<G, 2>(kcol, y, tmp); break;
                        case 3: stage_argument<G, 3>(kcol, y, tmp); break;
                        case 4: stage_argument<G, 4>(kcol, y, tmp); break;
#ifdef GRRT_DEFER_SUMS
                        default: stage_argument_final<G>(kcol, y, tmp, y_new, acc_e); break;
#else
                        default: stage_argument<G, 5>(kcol, y, tmp); break;
#endif
                    }
                    double d[8];
#if defined(GRRT_SINCOS_DELTA) && defined(GRRT_FAST_MATH)
                    if (GeomTraits<G>::chart != CH_CART) {
                        // sin, cos of the stage's polar angle theta + delta from those of the state's theta (computed
                        // once per attempt) and a short polynomial in the small increment delta = sum B k; 99.9 % of
                        // all increments are below 1/8, larger ones take the full routine
                        double st, ct;
                        if (s == 0) {
                            sincos_pi2(y[2], &sin0, &cos0);
                            st = sin0;
                            ct = cos0;
                        } else {
                            const double delta = tmp[2];
                            if (fabs(delta) <= 0.125) {
                                double sd, cd;
                                sincos_small(delta, &sd, &cd);
                                st = fma(sin0, cd, cos0 * sd);
                                ct = fma(cos0, cd, -(sin0 * sd));
                            } else {
                                sincos_cold(y[2] + delta, &st, &ct);   // out of line: keeps the straight-line stage code dense
                            }
                        }
                        geodesic_rhs_trig<G>(S, C, tmp, st, ct, d);
                    } else
#endif
                    geodesic_rhs<G>(S, C, tmp, d);
#if defined(GRRT_UNROLL_STAGES) && !defined(GRRT_FAST_MATH)
                    const double ch = rk_ch(s), ct = rk_ct(s);
#else
                    const double ch = c_rk_ch[s], ct = c_rk_ct[s];
#endif
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (Comp<G>::on(i)) {
                            const double k = h_cur * d[i];
                            if (s < 5 && Comp<G>::stored(i)) kst[Comp<G>::slot(i) * kTraceThreads] = k;
#ifdef GRRT_DEFER_SUMS
                            if (s < 5 && Comp<G>::stored(i)) continue;   // summed at the last stage (stage_argument_final)
#endif
#ifdef GRRT_FAST_MATH
                            // CH2 = CT2 = 0 (runge_kutta.rs:43-55): the exact build adds the zeros; skipping them is worth
                            // 3 % for Schwarzschild (C1 382 -> 370 ms) but costs KerrBL 7 % (register allocation at 96)
#ifdef GRRT_UNROLL_STAGES
                            if (s != 1) {
#else
                            if (G == GRRT_GEOM_KERR_BL || s != 1) {
#endif
                                y_new[i] = y_new[i] + ch * k;
                                acc_e[i] = acc_e[i] + ct * k;
                            }
#else
                            y_new[i] = y_new[i] + ch * k;
                            acc_e[i] = acc_e[i] + ct * k;
#endif
                        }
                    }
                    kst += NC * kTraceThreads;
                }
                // truncation error: nalgebra's 8-accumulator dot product, (0,4) (1,5) (2,6) (3,7)
                double e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = Comp<G>::on(i) ? acc_e[i] * acc_e[i] : 0.0;
                double res = (e[0] + e[4]);
                res = res + (e[1] + e[5]);
                res = res + (e[2] + e[6]);
                res = res + (e[3] + e[7]);
                const double err = sqrt(res);
                ++R.attempts;
                double h_proposed;
#ifdef GRRT_FAST_MATH
                // (epsilon/err)^(1/5) = q^(-1/5), q = err/epsilon: single-precision seed, then two steps of the
                // division-free Newton iteration z <- z (1.2 - 0.2 q z^5) (error 3e^2 per step: 5e-6 -> 8e-11 -> < 1 ulp).
                // Outside the seed's range (and for err = 0, inf, NaN) the exact path below decides.
                const double q = err * S.inv_epsilon;
                if (q > 1e-30 && q < 1e30) {
                    double z = (double)exp2f(-0.2f * __log2f((float)q));
                    const double mq = -0.2 * q;
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const double z2 = z * z;
                        z = z * fma(mq, z2 * z2 * z, 1.2);
                    }
                    h_proposed = 0.9 * h_cur * z;
                } else
#endif
#ifdef GRRT_FAST_MATH
                if (err > 0.0) h_proposed = 0.9 * h_cur * controller_pow_cold(S.epsilon, err);
#else
                if (err > 0.0) h_proposed = 0.9 * h_cur * pow_glibc(pow_tab, S.epsilon / err, 1.0 / 5.0);
#endif
                else h_proposed = h_cur * kHGrowthCap;
                h_proposed = rclamp(fmin(h_proposed, h_cur * kHGrowthCap), kHMin, kHMax);
                if (err > S.epsilon) {
                    if (h_cur <= kHMin) {
                        accepted = true;  // accept what we have at the minimum step
                        h = h_cur;
                    } else {
                        h_cur = rclamp(h_proposed / 2.0, kHMin, kHMax);
                        if (++R.retries >= kMaxRetryStep) {
                            finish = GRRT_STOP_ERROR;
                            ray_error = GRRT_RAYERR_RKF45_RETRIES;
                        }
                    }
                } else {
                    accepted = true;
#ifdef GRRT_FAST_MATH
                    h = (err * S.inv_epsilon < 1e-5) ? rclamp(h_cur * kHGrowthCap, kHMin, kHMax) : h_proposed;
#else
                    h = (err / S.epsilon < 1e-5) ? rclamp(h_cur * kHGrowthCap, kHMin, kHMax) : h_proposed;
#endif
                }
                // ---------------- accepted step: cheap stop tests; chords that may touch something become events ----------------
                if (accepted) {
                    // far from every object and from the celestial sphere the chord needs no Cartesian end points and
                    // the previous state (only read when a hit is shaded) is not kept; otherwise the step is handed to
                    // the event code below (chord tests, escape test), which keeps the attempt loop short
                    const bool need_geo = segment_needs_cartesian<G>(S, y[1], y[2], y_new[1], y_new[2]);
                    if (need_geo) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) s_yprev[i * kTraceThreads + tid] = y[i];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = y_new[i];
                    ++R.step;
                    R.retries = 0;
                    h_cur = rclamp(h, kHMin, kHMax);
                    geo_pending = need_geo;
                    if (!need_geo) {
                        R.prev_valid = 0;
                        const int st = should_stop_inside<G>(S, y, R.step);
                        if (st != GRRT_STOP_NONE) finish = st;
                        else if (R.step >= S.max_steps - 1) finish = GRRT_STOP_NONE;  // budget exhausted (integrator.rs:173)
                    }
                }
                event = geo_pending || finish >= 0;
            }
        }

        // ================= events =================
        // ---------------- accepted steps whose chord may touch an object or the celestial sphere ----------------
        V3 prev_cart = v3(0.0, 0.0, 0.0), cart = v3(0.0, 0.0, 0.0);
        if (geo_pending) {
            if (R.prev_valid) {
                prev_cart = v3(CTX(CX_PREV_X), CTX(CX_PREV_Y), CTX(CX_PREV_Z));
            } else {
                double yp[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) yp[i] = s_yprev[i * kTraceThreads + tid];
                prev_cart = position_cartesian<G>(S, yp);
            }
            cart = position_cartesian<G>(S, y);
            // geometric half of Objects::intersects (objects.rs:65-120): which objects become the nearest
            // hit when their turn comes; they are shaded, in this order, below
            double shortest = 1.7976931348623157e308;
            for (int k = 0; k < S.n_objects; ++k) {
                const DObject& o = S.obj[k];
                Hit hcur;
                bool hit;
                if (o.kind == GRRT_OBJ_SPHERE) hit = sphere_intersects(o, prev_cart, cart, hcur);
                else if (o.kind == GRRT_OBJ_DISC) hit = disc_intersects(o, prev_cart, cart, hcur);
                else hit = HAS_VOL ? vol_intersects(o, prev_cart, cart, hcur) : false;  // no such object without HAS_VOL
                if (!hit) continue;
                const double distance = norm3(hcur.point - prev_cart);
                if (!(distance < shortest)) continue;
                shortest = distance;
                cand |= 1u << k;
            }
            CTX(CX_PREV_X) = cart.x;
            CTX(CX_PREV_Y) = cart.y;
            CTX(CX_PREV_Z) = cart.z;
            R.prev_valid = 1;
            const int st = should_stop<G>(S, y, cart, R.step);
            if (st != GRRT_STOP_NONE) finish = st;
            else if (R.step >= S.max_steps - 1) finish = GRRT_STOP_NONE;  // budget exhausted (integrator.rs:173)
        }
        auto push_hit = [&](const Cie& hit_color) {
            if (R.n_hits < kMaxHits) {
                hits[R.n_hits++] = hit_color;
            } else {  // fold into the last slot: "older over newer" keeps the compositing order
                hits[kMaxHits - 1] = blend(hit_color, hits[kMaxHits - 1]);
                if (A.counters) atomicAdd(A.counters + CNT_OVERFLOW, 1ull);
            }
            double alpha = rclamp(hit_color.a, 0.0, 1.0);
            CTX(CX_OPACITY) = alpha + CTX(CX_OPACITY) * (1.0 - alpha);  // scene.rs:149-150
        };
        // ---------------- shading half of Objects::intersects for the candidates, in object order ----------------
        int vol_k = -1;
        if (cand) {
            double yp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) yp[i] = s_yprev[i * kTraceThreads + tid];
            // prev_cart / cart: the two end points the chord tests above used
            int best_k = -1;
            double best_temperature = 0.0, best_redshift = 0.0;
            Hit best;
            best.u = best.v = best.t = 0.0;
            best.point = best.dir = v3(0.0, 0.0, 0.0);
            bool shading_error = false;
            for (int k = 0; k < S.n_objects; ++k) {
                if (!((cand >> k) & 1u)) continue;
                const DObject& o = S.obj[k];
                Hit hcur;
                bool hit;
                if (o.kind == GRRT_OBJ_SPHERE) hit = sphere_intersects(o, prev_cart, cart, hcur);
                else if (o.kind == GRRT_OBJ_DISC) hit = disc_intersects(o, prev_cart, cart, hcur);
                else hit = HAS_VOL ? vol_intersects(o, prev_cart, cart, hcur) : false;  // no such object without HAS_VOL
                if (!hit) continue;
                // step_at_intersection (objects.rs:27-44): exact point, lerped momentum
                double p0[4], p1[4], xs[4], ps[4];
                momentum_from_state<G>(S, C, yp, p0);
                momentum_from_state<G>(S, C, y, p1);
                double s = 1.0 - hcur.t;
#pragma unroll
                for (int i = 0; i < 4; ++i) ps[i] = s * p0[i] + hcur.t * p1[i];
                cartesian_to_native<G>(S, hcur.point, xs);
                double vel[4];
                if (o.kind == GRRT_OBJ_SPHERE) {
                    stationary_velocity<G>(S, xs, vel);
                } else if (!circular_orbit_velocity<G>(S, xs, vel)) {
                    shading_error = true;
                    ray_error = GRRT_RAYERR_NO_CIRCULAR_ORBIT;
                    break;
                }
                double emitter_energy = inner_product<G>(S, xs, vel, ps);
                best_redshift = (GeomTraits<G>::sig0 * CTX(CX_FREQ_E)) / (GeomTraits<G>::sig0 * emitter_energy);
                if (o.kind == GRRT_OBJ_SPHERE) {
                    best_temperature = o.sphere_temperature;
                } else {
                    int terr = compute_temperature(o.temp, radial_coordinate_cart<G>(S, hcur.point), best_temperature);
                    if (terr) {
                        shading_error = true;
                        ray_error = terr;
                        break;
                    }
                }
                best = hcur;
                best_k = k;
            }
            if (shading_error) {
                finish = GRRT_STOP_ERROR;  // `?` in Objects::intersects: the whole pixel fails (raytracer.rs:230-239)
            } else if (best_k >= 0) {
                const DObject& o = S.obj[best_k];
                if (HAS_VOL && o.kind == GRRT_OBJ_VOLUMETRIC_DISC) {
                    vol_k = best_k;  // shaded below by the whole warp
                    CTX(CX_VOL_PX) = best.point.x;
                    CTX(CX_VOL_PY) = best.point.y;
                    CTX(CX_VOL_PZ) = best.point.z;
                    CTX(CX_VOL_DX) = best.dir.x;
                    CTX(CX_VOL_DY) = best.dir.y;
                    CTX(CX_VOL_DZ) = best.dir.z;
                } else {
                    push_hit(texture_color(S, o.tex, best.u, best.v, best_temperature, best_redshift));
                }
            }
        }

        // ---------------- warp-cooperative ray-marches for lanes with a VolumetricDisc hit ----------------
        if (HAS_VOL) {
            __syncwarp();  // the hit lanes' shared-memory hand-off is visible to the whole warp
            unsigned pending = __ballot_sync(kFull, vol_k >= 0);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                const int k = __shfl_sync(kFull, vol_k, src);
                const int src_tid = (tid & ~31) | src;
                V3 ro, rd, dir;
                ro.x = s_ctx[CX_VOL_PX * kTraceThreads + src_tid];
                ro.y = s_ctx[CX_VOL_PY * kTraceThreads + src_tid];
                ro.z = s_ctx[CX_VOL_PZ * kTraceThreads + src_tid];
                dir.x = s_ctx[CX_VOL_DX * kTraceThreads + src_tid];
                dir.y = s_ctx[CX_VOL_DY * kTraceThreads + src_tid];
                dir.z = s_ctx[CX_VOL_DZ * kTraceThreads + src_tid];
                rd = dir / norm3(dir);  // .normalize() (volumetric_disc.rs:589-594)
                const double fe = s_ctx[CX_FREQ_E * kTraceThreads + src_tid];
                const double fpt = s_ctx[CX_FREQ_PT * kTraceThreads + src_tid];
                const double fpp = s_ctx[CX_FREQ_PPHI * kTraceThreads + src_tid];
                int merr = 0;
                unsigned long long smp = 0;
                const DObject& o = S.obj[k];
                Cie c = vol_march_warp<G>(S, o, s_perm + o.perm_slot * 256, ro, rd, fe, fpt, fpp, &merr, &smp);
                if (lane == src) {
                    push_hit(c);  // a failed march yields (0,0,0,0) (volumetric_disc.rs:596-600)
                    if (A.counters) atomicAdd(A.counters + CNT_SAMPLES, smp);
                }
            }
        }

        if (finish >= 0) {
            Cie result = cie(0.0, 0.0, 0.0, 1.0);
            int ray_class = GRRT_CLASS_ESCAPED;
            if (finish != GRRT_STOP_ERROR) {
                // scene.rs:157-214
                Cie terminal = cie(0.0, 0.0, 0.0, 1.0);
                bool has_terminal = false;
                if (finish == GRRT_STOP_HORIZON || finish == GRRT_STOP_CLOSED_ORBIT) {
                    has_terminal = true;
                    ray_class = GRRT_CLASS_CAPTURED;
                } else if (finish == GRRT_STOP_CELESTIAL) {
                    // get_uv_coordinates (scene.rs:222-231) on the last stored step
                    double theta, phi;
                    if (GeomTraits<G>::chart == CH_CART) {
                        double r;
                        cartesian_to_spherical(y[1], y[2], y[3], r, theta, phi);
                    } else {
                        theta = rem_euclid(y[2], kPi);
                        phi = rem_euclid(y[3] + kPi, 2.0 * kPi) - kPi;
                    }
                    double u = (kPi + phi) / (2.0 * kPi);
                    double v = theta / kPi;
                    // compute_redshift (redshift.rs:31-34): static emitter at the last step
                    double vel[4], p[4];
                    stationary_velocity<G>(S, y, vel);
                    momentum_from_state<G>(S, C, y, p);
                    double emitter_energy = inner_product<G>(S, y, vel, p);
                    double redshift = (GeomTraits<G>::sig0 * CTX(CX_FREQ_E)) / (GeomTraits<G>::sig0 * emitter_energy);
                    terminal = texture_color(S, S.celestial, 1.0 - u, v, S.celestial_temperature, redshift);
                    has_terminal = true;
                    ray_class = GRRT_CLASS_ESCAPED;
                } else {
                    ray_class = GRRT_CLASS_CAPTURED;  // NaN or no terminal event
                }
                if (has_terminal) result = blend(result, terminal);
                for (int k = R.n_hits - 1; k >= 0; --k) result = blend(result, hits[k]);
                if (CTX(CX_OPACITY) >= S.hit_opacity_threshold) ray_class = GRRT_CLASS_HIT;
            }
            // write out
            const unsigned long long o = R.unit;
            if (A.spa == 0 && A.n_peers) {
                // multi-GPU frame: the finished pixel goes straight into every rank's frame over NVLink (33 B per
                // peer per ray), so no gather / scatter pass follows the kernel
                for (unsigned p = 0; p < A.n_peers; ++p) {
                    double* dst = A.xyza_p[p] + 4 * o;
                    dst[0] = result.x; dst[1] = result.y; dst[2] = result.z; dst[3] = result.a;
                    A.cls_p[p][o] = (uint8_t)ray_class;
                    A.stop_p[p][o] = (uint8_t)finish;
                    if (A.steps_p[p]) A.steps_p[p][o] = (uint32_t)R.step;
                }
                if (A.steps) A.steps[o] = (uint32_t)R.step;
            } else if (A.spa == 0) {
                double* dst = A.xyza + 4 * o;
                dst[0] = result.x; dst[1] = result.y; dst[2] = result.z; dst[3] = result.a;
                if (A.cls) A.cls[o] = (uint8_t)ray_class;
                if (A.stop) A.stop[o] = (uint8_t)finish;
                if (A.steps) A.steps[o] = (uint32_t)R.step;
            } else {
                double* dst = A.xyza + 4 * o;
                dst[0] = result.x; dst[1] = result.y; dst[2] = result.z; dst[3] = result.a;
                A.sub_valid[o] = finish != GRRT_STOP_ERROR ? 1 : 0;
            }
            if (A.counters) {  // once per ray: negligible next to the thousands of attempts a ray takes
                atomicAdd(A.counters + CNT_ATTEMPTS, (unsigned long long)R.attempts);
                atomicAdd(A.counters + CNT_ACCEPTED, (unsigned long long)R.step);
                atomicAdd(A.counters + CNT_RAYS, 1ull);
                atomicAdd(A.counters + CNT_STOP0 + finish, 1ull);
                if (finish == GRRT_STOP_ERROR) atomicAdd(A.counters + CNT_ERRORS, 1ull);
            }
            state = ST_IDLE;
        }
    }

#undef CTX
}

// ---- launch ---------------------------------------------------------------------------------------------
template <int G, bool V, int MB>
int launch_trace_t(const DScene& s, const TraceArgs& a, const TraceLaunch& l) {
    auto kern = trace_kernel<G, V, MB>;
    const size_t smem = trace_smem_bytes<G, V>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTraceThreads, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    unsigned long long want = (a.n_units + kTraceThreads - 1) / kTraceThreads;
    unsigned long long grid = (unsigned long long)l.num_sms * per_sm;  // persistent: one wave, a multiple of the SM count
    if (want < grid) grid = want;
    if (grid == 0) return 0;
    kern<<<(unsigned)grid, kTraceThreads, smem, l.stream>>>(s, a);
    return (int)cudaGetLastError();
}
}  // namespace GRRT_NS
