"""Pins the CPU oracle against the known-answer values the reference's own unit tests hold
for the geodesic path.  Each test names the reference test it restates."""
import ctypes as C
import math

import numpy as np
import pytest

from gr_raytracer_b200 import abi
from oracle import scene as osc
from helpers import PI, checker, reference_test_scene, oracle_config_scene

L = osc.lib()


def vec4():
    return (C.c_double * 4)()


def ip(geom, pos, v, w):
    return L.orc_inner_product(C.byref(geom), osc.darr(pos), osc.darr(v), osc.darr(w))


def cam_ray(cam, row, col, offset=None):
    pos, mom = vec4(), vec4()
    dx, dy = offset or (0.0, 0.0)
    L.orc_camera_ray(C.byref(cam), row, col, 1 if offset else 0, dx, dy, pos, mom)
    return list(pos), list(mom)


# ---- rendering/camera.rs ---------------------------------------------------------------

def test_camera_corner_directions_euclidean():
    """camera.rs:274-336: FOV pi/2, 11x11, camera at (1,0,0)."""
    g = osc.make_geometry(abi.GEOM_EUCLIDEAN)
    cam = osc.make_camera(g, [0, 1, 0, 0], [1, 0, 0, 0], PI / 2, 11, 11)
    corner, corner_z = 0.6853582554517135, 0.24610591900311507
    exp = {(0, 0): (corner_z, corner, -corner), (0, 10): (corner_z, -corner, -corner),
           (10, 0): (corner_z, corner, corner), (10, 10): (corner_z, -corner, corner)}
    for (r, c), e in exp.items():
        _, mom = cam_ray(cam, r, c)
        d = [mom[i] + [1, 0, 0, 0][i] for i in range(4)]  # direction = momentum + e_t
        assert d[0] == pytest.approx(0.0, abs=1e-12)
        assert d[1:] == pytest.approx(list(e), abs=1e-8)
        assert ip(g, [0, 1, 0, 0], d, d) == pytest.approx(-1.0, abs=1e-8)
    _, mom = cam_ray(cam, 5, 5)
    assert [mom[0] + 1, mom[1], mom[2], mom[3]] == pytest.approx([0, -1, 0, 0], abs=1e-8)


def test_camera_centre_offset_equals_base_ray():
    """camera.rs:339-364: get_ray_for_offset(.., 0.5, 0.5) == get_ray_for."""
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    cam = osc.make_camera(g, [0, 10, PI / 2, 0], [1 / math.sqrt(0.9), 0, 0, 0], PI / 4, 101, 101)
    assert cam_ray(cam, 17, 23) == cam_ray(cam, 17, 23, (0.5, 0.5))


@pytest.mark.parametrize("kind,pos,sig", [
    (abi.GEOM_SCHWARZSCHILD, [-10.0, 0.0, 2.0], 1.0),
    (abi.GEOM_KERR, [-10.0, 0.0, 2.0], -1.0),
    (abi.GEOM_KERR_BL, [-10.0, 0.0, 2.0], -1.0),
])
def test_camera_rays_null_and_past_directed(kind, pos, sig):
    """camera.rs:462-508 and kerr.rs:700-726: k.k = 0 and sig0 * g(u, p) < 0 for every pixel."""
    g = osc.make_geometry(kind, 1.0, 0.4 if kind != abi.GEOM_SCHWARZSCHILD else 0.0, 1e-5)
    p = osc.camera_position_native(g, pos)
    u = osc.resolve_camera_velocity(g, p)
    cam = osc.make_camera(g, p, u, PI / 2, 11, 11)
    for r in range(0, 11, 2):
        for c in range(0, 11, 5):
            _, mom = cam_ray(cam, r, c)
            assert abs(ip(g, p, mom, mom)) < 1e-8
            assert sig * ip(g, p, u, mom) < 0.0


def test_tetrads_orthonormal_all_geometries():
    """schwarzschild.rs:408-440, kerr.rs:561-587, kerr_bl.rs:1037-1094."""
    for kind, a in [(abi.GEOM_SCHWARZSCHILD, 0.0), (abi.GEOM_KERR, 0.499), (abi.GEOM_KERR_BL, 0.499),
                    (abi.GEOM_EUCLIDEAN, 0.0), (abi.GEOM_EUCLIDEAN_SPHERICAL, 0.0)]:
        g = osc.make_geometry(kind, 1.0, a, 1e-4)
        p = osc.camera_position_native(g, [3.0, -2.0, 1.5])
        t = (C.c_double * 16)()
        L.orc_tetrad_at(C.byref(g), osc.darr(p), t)
        e = [list(t[4 * i:4 * i + 4]) for i in range(4)]
        s0 = -1.0 if kind in (abi.GEOM_KERR, abi.GEOM_KERR_BL) else 1.0
        sig = [s0, -s0, -s0, -s0]
        for i in range(4):
            for j in range(4):
                assert ip(g, p, e[i], e[j]) == pytest.approx(sig[i] if i == j else 0.0, abs=1e-10)


# ---- rendering/scene.rs end-to-end known answers ------------------------------------------

SPHERE_COLOR_2 = [0.4124564, 0.2126729, 0.0193339, 1.0]
CELESTIAL_SPHERE_COLOR_2 = [0.3575761, 0.7151522, 0.119192, 1.0]


def test_color_of_ray_hits_sphere_euclidean():
    """scene.rs:416-438."""
    g = osc.make_geometry(abi.GEOM_EUCLIDEAN)
    cam = osc.make_camera(g, [0, 10, 0, 0], [1, 0, 0, 0], PI / 2, 11, 11)
    sc = reference_test_scene(g, cam, 2.0, 0.2, 0.3)
    r = osc.color_of_ray(sc, 5, 5)
    assert r["ray_class"] == abi.CLASS_HIT
    assert r["xyza"] == pytest.approx(SPHERE_COLOR_2, abs=1e-6)


def test_color_of_ray_hits_sphere_spherical():
    """scene.rs:441-477."""
    g = osc.make_geometry(abi.GEOM_EUCLIDEAN_SPHERICAL)
    p = osc.camera_position_native(g, [0, 0, -10])
    cam = osc.make_camera(g, p, [1, 0, 0, 0], PI / 2, 11, 11)
    sc = reference_test_scene(g, cam, 2.0, 0.2, 0.3)
    r = osc.color_of_ray(sc, 5, 5)
    assert r["xyza"] == pytest.approx([0.052562486896837575, 0.0271025410675224, 0.002463867369774764, 1.0],
                                      abs=1e-6)


@pytest.mark.parametrize("free_fall", [True, False])
def test_color_of_ray_hits_sphere_schwarzschild(free_fall):
    """scene.rs:480-538 (free-falling and static observer)."""
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    r, a = 10.0, 1.0 - 1.0 / 10.0
    vel = [-1.0 / a, -math.sqrt(1.0 / r), 0, 0] if free_fall else [-1.0 / math.sqrt(a), 0, 0, 0]
    cam = osc.make_camera(g, [0, r, PI / 2, 0], vel, PI / 2, 11, 11)
    sc = reference_test_scene(g, cam, 2.0, 3.0, 4.0)
    assert osc.color_of_ray(sc, 5, 5)["xyza"] == pytest.approx(SPHERE_COLOR_2, abs=1e-6)


def test_color_of_ray_misses_sphere():
    """scene.rs:541-602 (Euclidean and Schwarzschild r_s=2, theta=psi=pi/2)."""
    g = osc.make_geometry(abi.GEOM_EUCLIDEAN)
    cam = osc.make_camera(g, [0, 10, 0, 0], [1, 0, 0, 0], PI / 2, 11, 11)
    r = osc.color_of_ray(reference_test_scene(g, cam, 2.0, 0.2, 0.3), 0, 0)
    assert r["ray_class"] == abi.CLASS_ESCAPED
    assert r["xyza"] == pytest.approx(CELESTIAL_SPHERE_COLOR_2, abs=1e-6)

    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 2.0, 0.0, 1e-4)
    p = osc.camera_position_native(g, [10, 0, 0])
    a = 1.0 - 2.0 / p[1]
    cam = osc.make_camera(g, p, [1.0 / a, -math.sqrt(2.0 / p[1]), 0, 0], PI / 2, 11, 11, 0.0, PI / 2, PI / 2)
    r = osc.color_of_ray(reference_test_scene(g, cam, 2.0, 3.0, 4.0), 0, 0)
    assert r["xyza"] == pytest.approx(CELESTIAL_SPHERE_COLOR_2, abs=1e-6)


def test_color_of_ray_hits_horizon():
    """scene.rs:605-633."""
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    p = osc.camera_position_native(g, [-10, 0, 0])
    a = 1.0 - 1.0 / p[1]
    cam = osc.make_camera(g, p, [-1.0 / math.sqrt(a), 0, 0, 0], PI / 2, 11, 11, PI / 2, 0.0, PI / 2)
    r = osc.color_of_ray(reference_test_scene(g, cam, 0.5, 3.0, 4.0), 5, 5)
    assert r["ray_class"] == abi.CLASS_CAPTURED and r["stop"] == abi.STOP_HORIZON
    assert r["xyza"] == [0.0, 0.0, 0.0, 1.0]


def test_intersects_with_disk():
    """scene.rs:636-666."""
    g = osc.make_geometry(abi.GEOM_EUCLIDEAN)
    cam = osc.make_camera(g, [0, 7, 0, 0.8], [1, 0, 0, 0], PI / 4, 101, 101)
    r = osc.color_of_ray(reference_test_scene(g, cam, 1.0, 2.0, 7.0), 0, 51)
    assert r["xyza"] == pytest.approx([0.022994536463607135, 0.009197814585442854, 0.12110455021248553, 1.0],
                                      abs=1e-6)


# ---- rendering/runge_kutta.rs, integrator.rs -------------------------------------------------

def test_rkf45_simple_equation():
    """runge_kutta.rs:214-239: y'' = 2, y(0) = 1, y'(0) = 2."""
    out = (C.c_double * 3)()
    for t_end in (25.0, 50.0):
        L.orc_rkf45_simple_equation(t_end, 1e-10, 1e-7, out)
        t = out[0]
        assert t > t_end
        assert out[1] == pytest.approx(t * t + 2 * t + 1, abs=1e-5)
        assert out[2] == pytest.approx(2 * t + 2, abs=1e-5)


def test_should_stop_ordering():
    """integrator.rs:277-330."""
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 2.0, 0.0, 1e-5)
    y = osc.darr([0.0, 200.0, 1.0, 0.0, 1.0, 1.0, 0.0, math.inf])
    assert L.orc_should_stop(C.byref(g), y, 1, 100, 100.0) == abi.STOP_CELESTIAL
    y = osc.darr([0.0, 10.0, 1.0, 0.0, 1.0, 1.0, 0.0, math.nan])
    assert L.orc_should_stop(C.byref(g), y, 1, 100, 100.0) == abi.STOP_NAN
    y = osc.darr([0.0, 2.0, 1.0, 0.0, 1.0, 1.0, 0.0, 0.0])
    assert L.orc_should_stop(C.byref(g), y, 1, 100, 100.0) == abi.STOP_HORIZON
    y = osc.darr([0.0, 9.0, 1.0, 0.0, 1.0, 1.0, 0.0, 0.0])
    assert L.orc_should_stop(C.byref(g), y, 99, 100, 100.0) == abi.STOP_CLOSED_ORBIT
    assert L.orc_should_stop(C.byref(g), y, 98, 100, 100.0) == abi.STOP_NONE


def _bare_scene(geom, max_steps, max_radius, step, eps):
    sc = osc.Scene()
    sc.c.geometry = geom
    sc.c.integration.max_steps, sc.c.integration.max_radius = max_steps, max_radius
    sc.c.integration.step_size, sc.c.integration.epsilon = step, eps
    return sc


def test_celestial_sphere_reachable_with_cli_default_budget():
    """schwarzschild.rs:875-939: radial and 1.001*b_crit rays reach r = 15000 within 20000 steps."""
    radius, r0 = 2.0, 18.0
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, radius, 0.0, 1e-5)
    sc = _bare_scene(g, 20000, 15000.0, 0.01, 0.00001)
    a0 = 1.0 - radius / r0
    _, stop, err = osc.trace_ray(sc, [0, r0, PI / 2, 0], [1.0, a0, 0, 0], cap=1)
    assert (stop, err) == (abi.STOP_CELESTIAL, 0)
    r_ph = 1.5 * radius
    b = r_ph / math.sqrt(1.0 - radius / r_ph) * 1.001
    p_r = -math.sqrt(max(1.0 - a0 * b * b / (r0 * r0), 0.0))
    _, stop, err = osc.trace_ray(sc, [0, r0, PI / 2, 0], [1.0 / a0, p_r, 0, b / (r0 * r0)], cap=1)
    assert (stop, err) == (abi.STOP_CELESTIAL, 0)


def test_schwarzschild_vs_binet_rk4():
    """schwarzschild.rs:604-708 idea: the integrated orbit obeys Binet's equation
    u'' + u = 1.5 r_s u^2 (u = 1/r, prime = d/dphi); L = 5 escapes, L = 2 is captured."""
    radius = 1.0
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, radius, 0.0, 1e-4)
    sc = _bare_scene(g, 30000, 10000.0, 0.001, 1e-12)
    r0 = 10.0
    a0 = 1.0 - radius / r0
    for ang_mom, expect in ((5.0, abi.STOP_CELESTIAL), (2.0, abi.STOP_HORIZON)):
        p_t = 1.0 / a0
        p_phi = ang_mom / (r0 * r0)
        p_r = -math.sqrt(max(1.0 - a0 * ang_mom * ang_mom / (r0 * r0), 0.0))
        steps, stop, err = osc.trace_ray(sc, [0, r0, PI / 2, 0], [p_t, p_r, 0, p_phi])
        assert (stop, err) == (expect, 0)
        # independent RK4 on Binet's equation in phi
        r, phi = steps["x"][:, 1], steps["x"][:, 3]
        u, du = 1.0 / r0, -p_r / (r0 * r0) / p_phi  # du/dphi = -(dr/dl)/(r^2 dphi/dl)
        f = lambda uu, dd: (dd, 1.5 * radius * uu * uu - uu)
        matched, k, cur_phi = 0, 1, 0.0
        h = 1e-4
        while k < len(phi) and cur_phi < phi[-1] and r[k] < 100.0:
            k1 = f(u, du); k2 = f(u + 0.5 * h * k1[0], du + 0.5 * h * k1[1])
            k3 = f(u + 0.5 * h * k2[0], du + 0.5 * h * k2[1]); k4 = f(u + h * k3[0], du + h * k3[1])
            u += h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
            du += h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
            cur_phi += h
            while k < len(phi) and phi[k] <= cur_phi:
                if abs(1.0 / u - r[k]) < 0.1:
                    matched += 1
                k += 1
        assert matched >= 180


def test_rotated_rays_give_rotated_trajectories():
    """kerr.rs:728-765 (Kerr with radius 0 == flat): 90-degree rotated rays, equal step counts."""
    g = osc.make_geometry(abi.GEOM_KERR, 0.0, 0.0, 1e-4)
    cam = osc.make_camera(g, [2.0, 1.0, 0.0, 0.0], [1.0, 0, 0, 0], PI / 2, 11, 11, 0.0, PI / 2, PI / 2)
    sc = reference_test_scene(g, cam, 1.0, 2.0, 7.0, eps=1e-5)
    a, _, _ = osc.trace_ray(sc, pixel=(5, 10))
    b, _, _ = osc.trace_ray(sc, pixel=(0, 5))
    assert len(a) == len(b)
    assert np.allclose(a["x"][:, 1], b["x"][:, 1], atol=1e-5)
    assert np.allclose(a["x"][:, 2], -b["x"][:, 3], atol=1e-5)
    assert np.allclose(a["x"][:, 3], b["x"][:, 2], atol=1e-5)


def test_ray_export_length():
    """cli/schwarzschild.rs:133-172: --max-steps 10 stores steps 0..9 (11 CSV lines + trailing newline)."""
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    p = osc.camera_position_native(g, [18.0, 0.0, 0.8])
    cam = osc.make_camera(g, p, osc.resolve_camera_velocity(g, p), PI / 4, 500, 500)
    sc = _bare_scene(g, 10, 15000.0, 0.01, 1e-5)
    sc.c.camera = cam
    steps, stop, err = osc.trace_ray(sc, pixel=(250, 250))
    assert len(steps) == 10 and list(steps["step"]) == list(range(10)) and stop == abi.STOP_NONE


# ---- geometry -------------------------------------------------------------------------------

def test_circular_orbit_and_isco():
    """kerr.rs:768-779, circular_orbit.rs:151-157."""
    out = (C.c_double * 2)()
    assert L.orc_circular_orbit(1.0, 0.0, 3.0, out) == 0
    # u = u_t d_t + u_phi (-y, x, 0) at (x,y,z) = (0,3,0)
    assert out[0] == pytest.approx(1.414213562373095, abs=1e-8)
    assert -3.0 * out[1] == pytest.approx(-0.5773502691896257, abs=1e-8)
    assert L.orc_r_isco(1.0, 0.0) == pytest.approx(3.0, abs=1e-12)
    assert L.orc_r_isco(1.0, 0.499) == pytest.approx(0.6184853275875923, abs=1e-12)
    assert L.orc_circular_orbit(1.0, 0.0, 1.4, out) == 2  # NoCircularOrbitPossible inside the photon orbit


def test_bl_to_cartesian_embedding():
    """point.rs:238-251."""
    g = osc.make_geometry(abi.GEOM_KERR_BL, 1.0, 0.5, 1e-4)
    out = vec4()
    L.orc_step_to_cartesian(C.byref(g), osc.darr([0.0, 5.0, 1.2, 0.8]), out)
    assert list(out)[1:] == pytest.approx([2.91248746519832302226, 3.66769851865865170737, 1.81178877238336810684],
                                          abs=1e-10)
    # and back (spherical_coordinates_helper.rs:44-61)
    back = osc.camera_position_native(g, list(out)[1:])
    assert back[1:] == pytest.approx([5.0, 1.2, 0.8], abs=1e-12)


def test_kerr_bl_rhs_structure():
    """kerr_bl.rs:1576-1631 at (r=5, theta=1.2, E=1, L=3, Q=1)."""
    # The solver constants come from the ray, so pick them through a ray that yields E, L_z, Q and
    # check the RHS formulas for those values.
    g = osc.make_geometry(abi.GEOM_KERR_BL, 1.0, 0.5, 1e-4)
    sc = _bare_scene(g, 10, 100.0, 0.01, 1e-5)
    pos, mom = [0.0, 5.0, 1.2, 0.0], [1.3, 0.1, -0.05, 0.11]
    consts, y0 = (C.c_double * 3)(), (C.c_double * 8)()
    L.orc_solver_constants(sc.ptr, osc.darr(pos), osc.darr(mom), consts, y0)
    e, l_z, q = list(consts)
    r, theta, v_r, v_theta, a, r_s = 5.0, 1.2, 0.1, -0.05, 0.5, 1.0
    y = [0.0, r, theta, 0.0, v_r, v_theta, 0.0, 0.0]
    rhs = (C.c_double * 8)()
    L.orc_geodesic_rhs(sc.ptr, osc.darr(pos), osc.darr(mom), osc.darr(y), rhs)
    delta = r * r - r_s * r + a * a
    p_r = (r * r + a * a) * e - a * l_z
    sin2 = math.sin(theta) ** 2
    assert rhs[0] == pytest.approx((r * r + a * a) / delta * p_r + a * (l_z - a * e * sin2), abs=1e-12)
    assert rhs[1] == v_r and rhs[2] == v_theta
    assert rhs[3] == pytest.approx(a / delta * p_r + l_z / sin2 - a * e, abs=1e-12)
    carter = (l_z - a * e) ** 2 + q
    assert rhs[4] == pytest.approx((4 * r * e * p_r - (2 * r - r_s) * carter) / 2, abs=1e-12)
    ct, st = math.cos(theta), math.sin(theta)
    assert rhs[5] == pytest.approx((-2 * a * a * e * e * ct * st + 2 * l_z * l_z * ct / st ** 3) / 2, abs=1e-12)
    assert rhs[6] == 0.0 and rhs[7] == 0.0


def test_kerr_bl_conservation_and_null_condition():
    """kerr_bl.rs:1216-1344: E, L_z, Q drift < 1e-4 and |k.k| < 1e-4 along a trajectory (a=0.4, eps=1e-6)."""
    g = osc.make_geometry(abi.GEOM_KERR_BL, 1.0, 0.4, 1e-5)
    p = osc.camera_position_native(g, [-10.0, 0.0, 2.0])
    cam = osc.make_camera(g, p, osc.resolve_camera_velocity(g, p), PI / 2, 11, 11)
    sc = reference_test_scene(g, cam, 1.0, 2.0, 7.0, eps=1e-6)
    steps, stop, err = osc.trace_ray(sc, pixel=(3, 7))
    assert err == 0 and len(steps) > 10
    consts0, y0 = (C.c_double * 3)(), (C.c_double * 8)()
    L.orc_solver_constants(sc.ptr, osc.darr(steps["x"][0]), osc.darr(steps["p"][0]), consts0, y0)
    for k in range(1, len(steps), max(1, len(steps) // 200)):
        x, pp = list(steps["x"][k]), list(steps["p"][k])
        assert abs(ip(g, x, pp, pp)) < 1e-4
        c = (C.c_double * 3)()
        L.orc_solver_constants(sc.ptr, osc.darr(x), osc.darr(pp), c, y0)
        for i in range(3):
            ref = consts0[i]
            drift = abs(c[i] - ref) / abs(ref) if abs(ref) > 1e-12 else abs(c[i] - ref)
            assert drift < 1e-4


def test_kerr_bl_matches_kerr_schild_and_schwarzschild():
    """kerr_bl.rs:1127-1213, 1347-1450: same physical ray, three charts, same fate and close end points."""
    ends = {}
    for kind in (abi.GEOM_KERR, abi.GEOM_KERR_BL):
        g = osc.make_geometry(kind, 1.0, 0.4, 1e-5)
        p = osc.camera_position_native(g, [-10.0, 0.0, 2.0])
        cam = osc.make_camera(g, p, osc.resolve_camera_velocity(g, p), PI / 2, 11, 11,
                              theta=(1.5707963 if kind == abi.GEOM_KERR else -3.14159265), psi=(-1.5707963 if kind == abi.GEOM_KERR else 0.0))
        sc = reference_test_scene(g, cam, 0.1, 0.2, 0.3, eps=1e-6)
        steps, stop, err = osc.trace_ray(sc, pixel=(5, 5))
        out = vec4()
        L.orc_step_to_cartesian(C.byref(g), osc.darr(steps["x"][-1]), out)
        ends[kind] = (stop, np.array(list(out)[1:]))
    # both central rays fall into the hole
    assert ends[abi.GEOM_KERR][0] in (abi.STOP_HORIZON, abi.STOP_CLOSED_ORBIT)
    assert ends[abi.GEOM_KERR_BL][0] == abi.STOP_HORIZON

    # a = 0: KerrBL reduces to Schwarzschild (same chart) -- end points within 100 on a 1e4 sphere
    res = {}
    for kind in (abi.GEOM_SCHWARZSCHILD, abi.GEOM_KERR_BL):
        g = osc.make_geometry(kind, 1.0, 0.0, 1e-5)
        p = osc.camera_position_native(g, [-10.0, 0.0, 2.0])
        cam = osc.make_camera(g, p, osc.resolve_camera_velocity(g, p), PI / 2, 11, 11, theta=-3.14159265)
        sc = reference_test_scene(g, cam, 0.1, 0.2, 0.3, eps=1e-6)
        steps, stop, err = osc.trace_ray(sc, pixel=(1, 8))
        out = vec4()
        L.orc_step_to_cartesian(C.byref(g), osc.darr(steps["x"][-1]), out)
        res[kind] = (stop, np.array(list(out)[1:]))
    assert res[abi.GEOM_SCHWARZSCHILD][0] == res[abi.GEOM_KERR_BL][0] == abi.STOP_CELESTIAL
    assert np.linalg.norm(res[abi.GEOM_SCHWARZSCHILD][1] - res[abi.GEOM_KERR_BL][1]) < 100.0


# ---- redshift ---------------------------------------------------------------------------------

def test_gravitational_redshift_static_observers():
    """redshift.rs:279-397: static camera at r_cam looking at the static celestial emitter:
    g = sqrt(a_em / a_cam) with a = 1 - r_s / r (a_em -> 1 at r = 1e4)."""
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    r_cam = 10.0
    cam = osc.make_camera(g, [0, r_cam, PI / 2, 0], [1 / math.sqrt(1 - 1 / r_cam), 0, 0, 0], PI / 2, 11, 11)
    sc = reference_test_scene(g, cam, 0.1, 3.0, 4.0, eps=1e-10)
    sc.c.celestial = checker(1.0, 1.0, 1.0, (255, 255, 255), (255, 255, 255))  # colour * g^1
    r = osc.color_of_ray(sc, 0, 0)
    assert r["stop"] == abi.STOP_CELESTIAL
    white_y = 0.2126729 + 0.7151522 + 0.0721750
    g_meas = r["xyza"][1] / white_y
    steps, _, _ = osc.trace_ray(sc, pixel=(0, 0))
    a_em = 1.0 - 1.0 / steps["x"][-1][1]
    assert g_meas == pytest.approx(math.sqrt(a_em / (1 - 1 / r_cam)), rel=1e-6)


# ---- colour / texture / LUTs ------------------------------------------------------------------

def test_black_body_to_srgb():
    """black_body_radiation.rs:63-73."""
    for temp, exp in ((1000.0, (255, 60, 0)), (10000.0, (137, 146, 172))):
        xyz = vec4()
        L.orc_blackbody_xyz(temp, 1.0, xyz)
        exposure = 1.0 / (xyz[0] + xyz[1] + xyz[2])
        out = (C.c_uint8 * 3)()
        L.orc_xyz_to_srgb(xyz, exposure, out)
        assert tuple(out) == exp


def test_srgb_roundtrip_and_blend():
    """color.rs:339-404."""
    xyz = vec4()
    L.orc_srgb_to_xyz(255, 42, 10, xyz)
    out = (C.c_uint8 * 3)()
    L.orc_xyz_to_srgb(xyz, 1.0, out)
    assert tuple(out) == (255, 42, 10)
    res = vec4()
    L.orc_blend(osc.darr([0.2, 0.4, 0.6, 1.0]), osc.darr([0.8, 0.1, 0.3, 0.0]), res)
    assert list(res) == pytest.approx([0.2, 0.4, 0.6, 1.0], abs=2.3e-16)
    L.orc_blend(osc.darr([0.2, 0.4, 0.6, 0.0]), osc.darr([0.8, 0.1, 0.3, 0.0]), res)
    assert list(res) == [0.0, 0.0, 0.0, 0.0]
    L.orc_blend(osc.darr([0.2, 0.4, 0.6, 1.0]), osc.darr([0.6, 0.4, 0.2, 0.5]), res)
    assert list(res) == pytest.approx([0.4, 0.4, 0.4, 1.0], abs=2.3e-16)


def test_bitmap_bilinear_2x2():
    """texture.rs:332-506: 2x2 RGBA checker, exact equality."""
    img = np.zeros((2, 2, 4), dtype=np.uint8)
    for y in range(2):
        for x in range(2):
            img[y, x] = (255, 0, 0, 128) if (x + y) % 2 == 0 else (0, 0, 255, 128)
    t = abi.GrrtTexture()
    t.kind, t.width, t.height, t.beaming_exponent = abi.TEX_BITMAP, 2, 2, 3.0
    t.rgba = osc.u8ptr(img)
    sc = osc.Scene()
    red, blue = vec4(), vec4()
    L.orc_srgb_to_xyz(255, 0, 0, red)
    L.orc_srgb_to_xyz(0, 0, 255, blue)

    def col(u, v):
        out = vec4()
        L.orc_texture_color(sc.ptr, C.byref(t), u, v, 0.0, 1.0, out)
        return list(out)
    assert col(0.0, 0.0) == [red[0], red[1], red[2], 128.0 / 255.0]
    assert col(0.999, 0.999) == [red[0], red[1], red[2], 128.0 / 255.0]
    assert col(0.0, 0.999) == [blue[0], blue[1], blue[2], 128.0 / 255.0]
    assert col(0.999, 0.0) == [blue[0], blue[1], blue[2], 128.0 / 255.0]
    mid = col(0.25, 0.25)
    assert mid == [(red[i] + blue[i]) / 2.0 for i in range(3)] + [128.0 / 255.0]


def test_blackbody_lut_matches_direct_integration():
    """texture.rs:476-489: LUT vs direct Planck integration within 2 %."""
    sc = osc.Scene()
    sc.blackbody_lut()
    t = abi.GrrtTexture()
    t.kind, t.beaming_exponent = abi.TEX_BLACKBODY, 0.0
    for temp in (1000.0, 5000.0, 10000.0, 100000.0):
        for z in (0.5, 1.0, 2.0):
            lut, direct = vec4(), vec4()
            L.orc_texture_color(sc.ptr, C.byref(t), 0.0, 0.0, temp, z, lut)
            L.orc_blackbody_xyz(temp, z, direct)
            for i in range(3):
                assert lut[i] == pytest.approx(direct[i], rel=0.02, abs=1e-14)
    # artistic exponent stacks on top (texture.rs:446-473)
    t4 = abi.GrrtTexture()
    t4.kind, t4.beaming_exponent = abi.TEX_BLACKBODY, 4.0
    a, b = vec4(), vec4()
    L.orc_texture_color(sc.ptr, C.byref(t), 0.0, 0.0, 6000.0, 1.5, a)
    L.orc_texture_color(sc.ptr, C.byref(t4), 0.0, 0.0, 6000.0, 1.5, b)
    assert b[0] == pytest.approx(a[0] * 1.5 ** 4, rel=1e-12)


def test_temperature_lut_errors_and_interpolation():
    """temperature.rs:197-253."""
    sc = osc.Scene()
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    tc = sc.temperature(g, 2000.0, 5.0)
    assert tc.r_isco == pytest.approx(3.0, abs=1e-12)
    out = C.c_double()
    assert L.orc_compute_temperature(C.byref(tc), math.nan, C.byref(out)) == 4   # NonFiniteRadius
    assert L.orc_compute_temperature(C.byref(tc), 2.9, C.byref(out)) == 3        # BelowRISCO
    assert L.orc_compute_temperature(C.byref(tc), 3.0, C.byref(out)) == 0 and out.value == tc.lut_t[0]
    assert L.orc_compute_temperature(C.byref(tc), 50.0, C.byref(out)) == 0 and out.value == tc.lut_t[999]
    temps = np.array([tc.lut_t[i] for i in range(1000)])
    # calibration: the hottest of the 10 probe radii is `temperature` (temperature.rs:78-101)
    assert temps.max() == pytest.approx(2000.0, rel=0.01)
    r_mid = 0.5 * (tc.lut_r[10] + tc.lut_r[11])
    assert L.orc_compute_temperature(C.byref(tc), r_mid, C.byref(out)) == 0
    assert out.value == pytest.approx(0.5 * (tc.lut_t[10] + tc.lut_t[11]), rel=1e-12)


# ---- rendering/raytracer.rs adaptive sampling ---------------------------------------------------

def _ad(**kw):
    ad = osc._adaptive_from_config(kw, abi.default_cli_opts())
    return ad


def _pair(p, pc, q, qc, ad, floor):
    return bool(L.orc_should_supersample_pair(osc.darr(p), pc, osc.darr(q), qc, C.byref(ad), floor))


def test_mask_logic():
    """raytracer.rs:527-618."""
    ESC, CAP, HIT = abi.CLASS_ESCAPED, abi.CLASS_CAPTURED, abi.CLASS_HIT
    s = lambda y, alpha: [0.0, y, 0.0, alpha]
    assert _pair(s(0, 1), ESC, s(0, 1), CAP, _ad(), 100.0) and _pair(s(0, 1), CAP, s(0, 1), ESC, _ad(), 100.0)
    zero = _ad(luminance_contrast_threshold=0.0, opacity_contrast_threshold=0.0)
    assert not _pair(s(1, 0), ESC, s(100, 1), ESC, zero, 0.0)
    c02 = _ad(luminance_contrast_threshold=0.2, opacity_contrast_threshold=0.2)
    assert _pair(s(2, 1), HIT, s(1, 1), HIT, c02, 1.0)
    assert _pair(s(2, 0.6), HIT, s(2, 0.9), HIT, c02, 1.0)
    assert not _pair(s(1, 0), HIT, s(0, 1), HIT, zero, 1.0)
    # luminance_contrast(0, 1e-4) = 0.5: just above / below a 0.5 threshold
    assert _pair(s(0, 1), HIT, s(1e-4, 1), HIT, _ad(luminance_contrast_threshold=0.499, opacity_contrast_threshold=1.0), 0.0)
    assert not _pair(s(0, 1), HIT, s(1e-4, 1), HIT, _ad(luminance_contrast_threshold=0.5, opacity_contrast_threshold=1.0), 0.0)


def test_stratified_offsets_stay_in_their_cells():
    n = 4
    dx, dy = C.c_double(), C.c_double()
    for sr in range(n):
        for sc_ in range(n):
            L.orc_stratified_sample_offset(17, 23, sr, sc_, n, C.byref(dx), C.byref(dy))
            assert sc_ / n <= dx.value < (sc_ + 1) / n and sr / n <= dy.value < (sr + 1) / n
            first = (dx.value, dy.value)
            L.orc_stratified_sample_offset(17, 23, sr, sc_, n, C.byref(dx), C.byref(dy))
            assert first == (dx.value, dy.value)
    # splitmix64 finaliser known value: mix64 chain is a pure function of (row, col, k)
    z = L.orc_hash_pixel_samples(0, 0, 0)
    assert 0.0 <= z < 1.0


# ---- VolumetricDisc -------------------------------------------------------------------------------

def test_volumetric_density_support():
    """volumetric_disc.rs:693-737: density > 0 inside the slab, == 0 outside radii / far above."""
    sc = oracle_config_scene("C5", width=8, height=8)
    o = sc.c.objects[0]
    dens = lambda p: L.orc_volumetric_density(C.byref(o), osc.darr(p))
    inside = [dens([r * math.cos(ph), r * math.sin(ph), 0.01]) for r in (5.0, 7.0, 9.0) for ph in (0.1, 1.3, 2.9, 4.4)]
    assert max(inside) > 0.0 and min(inside) >= 0.0
    assert dens([3.0, 0.0, 0.0]) == 0.0 and dens([12.0, 0.0, 0.0]) == 0.0 and dens([6.0, 0.0, 1.0]) == 0.0


def test_perlin_restatement_properties():
    """noise 0.9.0 is not vendored: only properties can be checked (zero at lattice points, range,
    permutation table is a permutation, determinism per seed)."""
    perm = (C.c_uint8 * 256)()
    L.orc_perlin_permutation(1, perm)
    assert sorted(perm) == list(range(256))
    perm2 = (C.c_uint8 * 256)()
    L.orc_perlin_permutation(2, perm2)
    assert list(perm) != list(perm2)
    assert L.orc_perlin(perm, 3.0, -2.0, 7.0) == 0.0
    rng = np.random.default_rng(0)
    vals = [L.orc_perlin(perm, *rng.uniform(-50, 50, 3)) for _ in range(2000)]
    assert -1.0 <= min(vals) < -0.3 and 0.3 < max(vals) <= 1.0


# ---- more of the reference's unit tests (metric, Killing decomposition, Doppler / Luminet, object geometry) --------

def test_kerr_schild_contravariant_metric_is_the_inverse():
    """kerr.rs:528-544 (test_metric_contravariant_matches_inverse): tolerance 1e-10 at the three listed points."""
    for x, y, z in [(3.0, -4.0, 1.5), (-7.5, 2.0, 0.3), (10.0, 1.0, -2.5)]:
        cov, contra = (C.c_double * 16)(), (C.c_double * 16)()
        L.orc_ks_metrics(1.0, 0.5, x, y, z, cov, contra)
        g = np.array(list(cov)).reshape(4, 4)
        gi = np.array(list(contra)).reshape(4, 4)
        assert np.allclose(g, g.T, atol=0) and np.abs(gi - np.linalg.inv(g)).max() < 1e-10


def _zamo(geom, pos):
    out = vec4()
    assert L.orc_resolve_camera_velocity(C.byref(geom), osc.darr(pos), abi.CAMVEL_ZAMO, osc.darr([0.0] * 4), out) == 0
    return list(out)


def _axial(geom, pos):
    out = vec4()
    L.orc_axial_killing_vector(C.byref(geom), osc.darr(pos), out)
    return list(out)


def test_zamo_properties_across_charts():
    """circular_orbit.rs:183-243: normalised, zero angular momentum, reduces to the static observer at a = 0."""
    a = 0.499
    for geom, pos in ((osc.make_geometry(abi.GEOM_KERR, 1.0, a, 1e-4), [0.0, 5.0, a, 0.0]),
                      (osc.make_geometry(abi.GEOM_KERR_BL, 1.0, a, 1e-4), [0.0, 5.0, PI / 2, 0.0])):
        u, ax = _zamo(geom, pos), _axial(geom, pos)
        assert ip(geom, pos, u, u) == pytest.approx(-1.0, abs=1e-9)
        assert ip(geom, pos, u, ax) == pytest.approx(0.0, abs=1e-9)
    gs = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    pos_s = [0.0, 5.0, 1.1, 0.3]
    stat = vec4()
    assert L.orc_resolve_camera_velocity(C.byref(gs), osc.darr(pos_s), abi.CAMVEL_STATIC, osc.darr([0.0] * 4), stat) == 0
    assert _zamo(gs, pos_s) == pytest.approx(list(stat), abs=1e-15)


@pytest.mark.parametrize("kind,a,pos,probes,time_sign", [
    (abi.GEOM_SCHWARZSCHILD, 0.0, [0.0, 6.0, PI / 2, 1.1], [[1.0, -0.4, 0.02, 0.05], [-1.3, 0.2, 0.0, -0.08]], 1.0),
    (abi.GEOM_KERR, 0.499, [0.0, 3.0, -4.0, 0.0], [[1.0, 0.3, -0.2, 0.1], [-0.8, -0.5, 0.4, 0.0]], -1.0),
    (abi.GEOM_KERR_BL, 0.3, [0.0, 5.0, PI / 2, 0.7], [[1.0, -0.4, 0.02, 0.05], [-1.2, 0.1, -0.03, 0.09]], -1.0),
])
def test_killing_decomposition_identities(kind, a, pos, probes, time_sign):
    """circular_orbit.rs:245-319: u = u_t d_t + u_phi xi_axial, u.u = signature[0], u.p = u_t p_t + u_phi p_phi."""
    geom = osc.make_geometry(kind, 1.0, a, 1e-4)
    u = vec4()
    assert L.orc_circular_orbit_velocity(C.byref(geom), osc.darr(pos), u) == 0
    u = list(u)
    ax = _axial(geom, pos)
    e_t = [1.0, 0.0, 0.0, 0.0]
    # the Killing coefficients: circular_orbit.rs:39-136 works with the equatorial BL radius of the point
    r_bl = pos[1] if kind != abi.GEOM_KERR else osc.camera_position_native(osc.make_geometry(abi.GEOM_KERR_BL, 1.0, a, 1e-4),
                                                                          pos[1:])[1]
    coeff = (C.c_double * 2)()
    assert L.orc_circular_orbit(1.0, a, r_bl, coeff) == 0
    assembled = [coeff[0] * e_t[i] + coeff[1] * ax[i] for i in range(4)]
    assert u == pytest.approx(assembled, abs=1e-12)
    assert ip(geom, pos, u, u) == pytest.approx(time_sign, abs=1e-10)
    for p in probes:
        direct = ip(geom, pos, u, p)
        assert direct == pytest.approx(coeff[0] * ip(geom, pos, e_t, p) + coeff[1] * ip(geom, pos, ax, p), abs=1e-10)


def _redshift(geom, pos, observer, emitter, momentum, sig0):
    """RedshiftComputer::compute_redshift_from_energies (redshift.rs:35-37, 69-73)."""
    return (sig0 * ip(geom, pos, observer, momentum)) / (sig0 * ip(geom, pos, emitter, momentum))


def test_doppler_formulas_with_hand_built_momentum():
    """redshift.rs:116-172: approaching / receding / transverse / at rest, and invariance under p -> -p."""
    g = osc.make_geometry(abi.GEOM_EUCLIDEAN, 0.0, 0.0, 0.0)
    pos, p, obs = [0.0] * 4, [-1.0, -1.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]
    v = 0.5
    gam = 1.0 / math.sqrt(1.0 - v * v)
    cases = {(gam, gam * v, 0.0, 0.0): 1.0 / (gam * (1.0 - v)), (gam, -gam * v, 0.0, 0.0): 1.0 / (gam * (1.0 + v)),
             (gam, 0.0, gam * v, 0.0): 1.0 / gam, (1.0, 0.0, 0.0, 0.0): 1.0}
    for em, want in cases.items():
        assert _redshift(g, pos, obs, list(em), p, -1.0) == pytest.approx(want, abs=1e-12)
    em = [gam, gam * v, 0.0, 0.0]
    assert _redshift(g, pos, obs, em, p, -1.0) == pytest.approx(_redshift(g, pos, obs, em, [-c for c in p], -1.0), abs=1e-15)


def test_disc_redshift_matches_luminet_closed_form():
    """redshift.rs:183-243: 1 + z = (1 - 3M/r)^(-1/2) (1 + Omega L/E) for a circular-orbit emitter in Schwarzschild."""
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    m = 0.5
    for r in (2.0, 3.0, 5.0, 10.0):
        a = 1.0 - 1.0 / r
        omega = math.sqrt(m / r ** 3)
        u_t = 1.0 / math.sqrt(1.0 - 3.0 * m / r)
        for phi in (0.3, 2.0, 4.5):
            pos = [0.0, r, PI / 2, phi]
            u = vec4()
            assert L.orc_circular_orbit_velocity(C.byref(g), osc.darr(pos), u) == 0
            u = list(u)
            for p_r, p_phi_hat in ((-0.7, 0.4), (-0.2, -0.9), (0.5, 0.6), (0.0, 1.0)):
                p_phi = p_phi_hat / r
                spatial = p_r * p_r / a + r * r * p_phi * p_phi
                p = [-math.sqrt(spatial / a), p_r, 0.0, p_phi]
                assert ip(g, pos, p, p) == pytest.approx(0.0, abs=1e-12)
                e_c, l_c = a * p[0], -r * r * p_phi
                emitter_energy = ip(g, pos, u, p)
                assert emitter_energy == pytest.approx(u_t * (e_c + omega * l_c), abs=1e-10)
                assert e_c / emitter_energy == pytest.approx(math.sqrt(1.0 - 3.0 * m / r) / (1.0 + omega * l_c / e_c), abs=1e-10)


def test_gravitational_redshift_analytic_without_integration():
    """redshift.rs:253-300: radial null ray, static emitter at r = 3 seen by a static camera at r = 10:
    g = sqrt(a_em / a_cam)."""
    g = osc.make_geometry(abi.GEOM_SCHWARZSCHILD, 1.0, 0.0, 1e-4)
    r_cam, r_em, p_t = 10.0, 3.0, -1.0
    a_cam, a_em = 1.0 - 1.0 / r_cam, 1.0 - 1.0 / r_em
    pos_c, pos_e = [0.0, r_cam, PI / 2, 0.0], [0.0, r_em, PI / 2, 0.0]
    p_c, p_e = [p_t / a_cam, p_t, 0.0, 0.0], [p_t / a_em, p_t, 0.0, 0.0]
    assert ip(g, pos_c, p_c, p_c) == pytest.approx(0.0, abs=1e-12) and ip(g, pos_e, p_e, p_e) == pytest.approx(0.0, abs=1e-12)
    stat = lambda pos: [1.0 / math.sqrt(1.0 - 1.0 / pos[1]), 0.0, 0.0, 0.0]
    shift = ip(g, pos_c, stat(pos_c), p_c) / ip(g, pos_e, stat(pos_e), p_e)
    assert shift == pytest.approx(math.sqrt(a_em / a_cam), abs=1e-12)


def _sphere_object(x, y, z, radius=1.0):
    o = abi.GrrtObject()
    o.kind = abi.OBJ_SPHERE
    o.texture = checker(0.0, 2.0, 2.0, (100, 0, 0), (0, 100, 0))
    o.radius = radius
    o.position[:] = [x, y, z]
    return o


def _hits(o, a, b):
    hit, t = (C.c_double * 3)(), C.c_double()
    rc = L.orc_object_intersects(C.byref(o), osc.darr(a), osc.darr(b), hit, C.byref(t))
    assert rc in (0, 1)
    return (list(hit), t.value) if rc else None


def test_sphere_intersections():
    """sphere.rs:188-246: a chord is a hit only if it crosses the surface; the hit point is in world space."""
    assert _hits(_sphere_object(0, 0, 0), [1.1, 0, 0], [0.9, 0, 0]) is not None
    assert _hits(_sphere_object(0, 0, 0), [1.1, 0, 0], [1.01, 0, 0]) is None
    assert _hits(_sphere_object(5, 0, 0), [6.1, 0, 0], [5.9, 0, 0]) is not None
    assert _hits(_sphere_object(5, 0, 0), [6.1, 0, 0], [6.01, 0, 0]) is None
    hit, _ = _hits(_sphere_object(0, 0, 20), [0, 0, 22.0], [0, 0, 19.5])
    assert hit[2] == pytest.approx(21.0, abs=1e-9)


def _reference_test_volumetric_disc(colour=None):
    """create_disc of volumetric_disc.rs:666-691 (Checker red/blue) or the FixedTextureMap variant (:741-760)."""
    o = abi.GrrtObject()
    o.kind = abi.OBJ_VOLUMETRIC_DISC
    o.texture = checker(3.0, 5.0, 5.0, (255, 0, 0), (0, 0, 255))
    if colour is not None:   # a texture that answers the same CIE value everywhere
        o.texture.c1[:] = colour
        o.texture.c2[:] = colour
        o.texture.beaming_exponent = 0.0
    o.temperature.kind = abi.TEMP_CONSTANT   # DummyTemperatureComputer: 1000 K at every radius
    o.temperature.constant = 1000.0
    o.inner_radius, o.outer_radius = 1.0, 3.0
    ax, e1, e2 = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
    L.orc_volumetric_basis(osc.darr([0.0, 0.0, 1.0]), ax, e1, e2)
    o.axis[:], o.e1[:], o.e2[:] = list(ax), list(e1), list(e2)
    o.num_octaves, o.vol_max_steps, o.vol_step_size, o.thickness = 4, 500, 0.01, 0.5
    o.density_multiplier, o.brightness_reference_temperature = 10.0, 1000.0
    o.absorption, o.scattering = 0.2, 0.2
    o.noise_scale[:] = [1.0, 1.0, 1.0]
    o.noise_offset = 1.0
    perm = (C.c_uint8 * 256)()
    L.orc_perlin_permutation(42, perm)
    o.perlin_perm[:] = list(perm)
    return o


def test_volumetric_disc_geometry_and_density():
    """volumetric_disc.rs:693-721: a radial chord through the gas hits, one above the caps misses; density is
    positive inside and exactly zero outside the radii / the slab."""
    o = _reference_test_volumetric_disc()
    assert _hits(o, [0.5, 0, 0], [1.5, 0, 0]) is not None
    assert _hits(o, [1.5, 0, 2.0], [2.5, 0, 2.0]) is None
    dens = lambda p: L.orc_volumetric_density(C.byref(o), osc.darr(p))
    assert dens([2.0, 0.0, 0.0]) > 0.0
    assert dens([0.2, 0.0, 0.0]) == 0.0 and dens([2.0, 0.0, 2.0]) == 0.0


def test_volumetric_raymarch_opacity_and_cached_exit_matches_legacy():
    """volumetric_disc.rs:723-786: 0 < alpha <= 1 for a march through the gas; the cached exit distance gives the
    result of the per-step exit test to 1e-10."""
    g = osc.make_geometry(abi.GEOM_EUCLIDEAN, 0.0, 0.0, 0.0)
    cam = osc.make_camera(g, [0.0, 10.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0], PI / 2, 3, 3)
    freq = osc.darr([1.0, 1.0, 0.0])   # unit_frequency(): observer_energy 1, p_t 1, p_phi 0
    for colour in (None, [2.0, 1.0, 0.5, 0.7]):
        sc = reference_test_scene(g, cam, 0.1, 3.0, 4.0)
        sc.c.objects[0] = _reference_test_volumetric_disc(colour)
        out = {}
        for cached in (1, 0):
            xyza, n = (C.c_double * 4)(), C.c_uint64()
            rc = L.orc_volumetric_raymarch(C.byref(sc.c), 0, osc.darr([2.0, 0.0, 0.0]), osc.darr([1.0, 0.0, 0.0]), freq,
                                           cached, xyza, C.byref(n))
            assert rc == 0
            out[cached] = list(xyza)
        assert 0.0 < out[1][3] <= 1.0
        assert out[1] == pytest.approx(out[0], abs=1e-10)
