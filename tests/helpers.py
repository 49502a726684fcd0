"""Shared fixtures: the reference's own test scene (src/rendering/scene.rs:234-370) rebuilt
from oracle-side primitives, and the five BASELINE.json configurations."""
import ctypes as C
import math
import os

import numpy as np

from gr_raytracer_b200 import abi
from oracle import scene as osc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PI = math.pi
CELESTIAL_SPHERE_RADIUS = 10000.0


def checker(beaming, w, h, c1, c2):
    """CheckerMapper::new, src/rendering/texture.rs:222-238."""
    t = abi.GrrtTexture()
    t.kind = abi.TEX_CHECKER
    t.beaming_exponent = beaming
    t.checker_width, t.checker_height = w, h
    for dst, col in ((t.c1, c1), (t.c2, c2)):
        out = (C.c_double * 4)()
        osc.lib().orc_srgb_to_xyz(col[0], col[1], col[2], out)
        dst[:] = list(out)
    return t


def reference_test_scene(geom, cam, sphere_r, disc_in, disc_out, eps=1e-12, max_steps=30000, step_size=0.001):
    """create_scene_with_camera, src/rendering/scene.rs:281-369."""
    sc = osc.Scene()
    sc.c.geometry = geom
    sc.c.camera = cam
    sc.c.integration.max_steps = max_steps
    sc.c.integration.max_radius = CELESTIAL_SPHERE_RADIUS
    sc.c.integration.step_size = step_size
    sc.c.integration.epsilon = eps
    sc.c.adaptive = osc._adaptive_from_config({}, abi.default_cli_opts())
    sc.c.celestial = checker(0.0, 100.0, 100.0, (0, 255, 0), (0, 100, 0))
    sc.c.celestial_temperature = 0.0
    o = abi.GrrtObject()
    o.kind = abi.OBJ_SPHERE
    o.texture = checker(0.0, 10.0, 10.0, (255, 0, 0), (100, 0, 0))
    o.radius = sphere_r
    o.position[:] = [0.0, 0.0, 0.0]
    o.sphere_temperature = 0.0
    sc.add_object(o)
    o = abi.GrrtObject()
    o.kind = abi.OBJ_DISC
    o.texture = checker(0.0, 200.0, 10.0, (0, 0, 255), (0, 0, 100))
    o.inner_radius, o.outer_radius = disc_in, disc_out
    o.temperature = sc.temperature(geom, 0.0, disc_out)
    sc.add_object(o)
    return sc


# BASELINE.json configs with the concrete command lines of SURVEY.md 8(d)
CONFIGS = {
    "C1": dict(toml="scenes/schwarzschild.toml", width=500, height=500, camera_position=[-16.0, 0.0, 3.5],
               theta=-3.142, psi=0.0, phi=0.0),
    "C2": dict(toml="scenes/kerr-bl.toml", width=1500, height=1500, camera_position=[-10.0, 0.0, -0.5],
               theta=-3.14159),
    "C3": dict(toml="scenes/kerr.toml", width=1500, height=1500, camera_position=[-10.0, 0.0, -0.5],
               theta=1.52, psi=-1.57, max_steps=1000000),
    "C4": dict(toml="scenes/kerr-volumetric-streaky.toml", width=2048, height=2048,
               camera_position=[-17.0, 0.0, 1.5], theta=1.659, psi=-1.57, phi=0.0, max_steps=1000000),
    "C5": dict(toml="scenes/schwarzschild-volumetric-dense.toml", width=8192, height=8192,
               camera_position=[-16.0, 0.0, 3.5], theta=-3.142),
}


def config_opts(name, **overrides):
    kw = {k: v for k, v in CONFIGS[name].items() if k != "toml"}
    kw.update(overrides)
    return abi.default_cli_opts(**kw)


def oracle_config_scene(name, **overrides):
    return osc.scene_from_toml(os.path.join(ROOT, CONFIGS[name]["toml"]), config_opts(name, **overrides), ROOT)


def rel_err_xyz(a, b):
    """Per-pixel relative error of linear XYZ as the north_star defines the gate:
    max_c |a_c - b_c| / max(|b|_inf, tiny)."""
    a = np.asarray(a, dtype=np.float64)[..., :3]
    b = np.asarray(b, dtype=np.float64)[..., :3]
    scale = np.maximum(np.max(np.abs(b), axis=-1), 1e-300)
    return np.max(np.abs(a - b), axis=-1) / scale
