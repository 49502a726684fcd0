"""Committed golden frames (tests/golden/*.npz, written by tests/golden/make_golden.py with the CPU oracle): the oracle
must still reproduce them (CPU), and the CUDA path must match them through the C ABI without the oracle being
involved at test time (GPU)."""
import os
import sys

import numpy as np
import pytest

from gr_raytracer_b200 import api
from oracle import scene as osc
from helpers import ROOT, CONFIGS, config_opts, oracle_config_scene, rel_err_xyz

GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)
from make_golden import CASES, RAY_PIXEL, RAY_STEPS  # noqa: E402


def load(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_frames(name):
    g = load(name)
    sc = oracle_config_scene(name, **CASES[name])
    frame = osc.render(sc)
    assert np.array_equal(frame["stop"], g["stop"]) and np.array_equal(frame["ray_class"], g["ray_class"])
    assert np.array_equal(frame["steps"], g["steps"])
    assert frame["stats"]["pixels_supersampled"] == int(g["pixels_supersampled"])
    # same compiler flags and libm give the same bits; the tolerance only allows for a different libm build
    assert np.allclose(frame["xyza"], g["xyza"], rtol=1e-9, atol=1e-300)
    if name in ("C1", "C2"):   # the 20 000-step trajectories of C3-C5 are covered by the frames
        steps, stop, _ = osc.trace_ray(sc, pixel=RAY_PIXEL[name])
        assert len(steps["x"]) == int(g["ray_len"]) and stop == int(g["ray_stop"])
        n = min(RAY_STEPS, len(steps["x"]))
        assert np.allclose(np.asarray(steps["x"][:n]), g["ray_x"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_matches_golden_frames(name):
    """north_star gate against the committed frames: XYZ within 1e-4 relative on >= 99.9 % of the pixels with the
    same termination reason and class; Kerr-Schild (C3, C4 geodesics) to the last bit of the step counts."""
    g = load(name)
    kw = CASES[name]
    scene = api.HostScene(os.path.join(ROOT, CONFIGS[name]["toml"]), config_opts(name, **kw), ROOT)
    with api.Context(0) as ctx:
        dev = ctx.render(scene)
        steps = ctx.trace_ray(scene, pixel=RAY_PIXEL[name])
    rel = rel_err_xyz(dev["xyza"], g["xyza"])
    black = np.max(np.abs(g["xyza"][..., :3]), axis=-1) == 0.0
    rel = np.where(black, np.max(np.abs(dev["xyza"][..., :3]), axis=-1), rel)
    good = (rel <= 1e-4) & (dev["stop"] == g["stop"]) & (dev["ray_class"] == g["ray_class"])
    print(f"[golden] {name}: {good.mean() * 100:.3f} % within 1e-4 & same fate, max rel err {rel.max():.3e}")
    assert good.mean() >= 0.999
    assert dev["stats"]["pixels_supersampled"] == int(g["pixels_supersampled"])
    if name in ("C3", "C4"):
        assert np.array_equal(dev["steps"], g["steps"])
    n = g["ray_x"].shape[0]
    got = np.asarray(steps[0]["x"][:n]) if isinstance(steps, tuple) else np.asarray(steps["x"][:n])
    assert np.allclose(got, g["ray_x"], rtol=1e-6, atol=1e-9)
