"""GPU edge cases of the render path (through the C ABI): empty and ragged sections, degenerate budgets,
scenes without objects, the exact-arithmetic build against the FMA build.  Each case is checked against the
CPU oracle on the same scene."""
import os

import numpy as np
import pytest

from gr_raytracer_b200 import abi, api
from oracle import scene as osc
from helpers import ROOT, CONFIGS, config_opts, oracle_config_scene
from test_gpu_parity import compare, both

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def _config(name, adaptive=None, drop_objects=False):
    import tomllib
    with open(os.path.join(ROOT, CONFIGS[name]["toml"]), "rb") as f:
        cfg = tomllib.load(f)
    if adaptive is not None:
        cfg["adaptive_sampling"] = {"enabled": adaptive}
    if drop_objects:
        cfg["objects"] = []
    return cfg


@pytest.mark.parametrize("name,w,h", [("C2", 37, 23), ("C1", 9, 5), ("C2", 1, 1), ("C1", 3, 70)])
def test_ragged_frame_sizes(ctx, name, w, h):
    """Frames that do not fill the kernel's 8x4 pixel tiles (and a single-pixel frame), adaptive pass included."""
    sc = oracle_config_scene(name, width=w, height=h)
    dev, ora = both(ctx, sc)
    compare(dev, ora, f"{name} {w}x{h}", min_fraction=1.0 if w * h < 200 else 0.999)
    assert dev["stats"]["rays_base"] == w * h
    assert dev["stats"]["pixels_supersampled"] == ora["stats"]["pixels_supersampled"]


def test_empty_inverted_and_overhanging_sections(ctx):
    """A zero-area section renders nothing and succeeds; an inverted one is an argument error; a section that
    overhangs the frame is legal in the reference (render_section only sizes its buffer by the section and asks
    the camera for whatever row/col it is given, raytracer.rs:177-244) and renders the pixels beyond the border."""
    sc = osc.scene_from_config(_config("C2", adaptive=False), config_opts("C2", width=16, height=12), ROOT)
    out = ctx.render(sc, 5, 3, 5, 9)
    assert out["xyza"].shape == (0, 6, 4) and out["stats"]["rays_base"] == 0
    out = ctx.render(sc, 2, 7, 9, 7)
    assert out["xyza"].shape == (7, 0, 4)
    with pytest.raises(api.GrrtError):
        ctx.render(sc, 6, 0, 5, 16)
    dev, ora = both(ctx, sc, from_row=8, from_col=10, to_row=15, to_col=21)
    compare(dev, ora, "C2 section overhanging the 16x12 frame", min_fraction=1.0)


def test_single_row_and_single_column_sections(ctx):
    sc = osc.scene_from_config(_config("C1", adaptive=False), config_opts("C1", width=24, height=20), ROOT)
    full = ctx.render(sc)
    row = ctx.render(sc, 11, 0, 12, 24)
    col = ctx.render(sc, 0, 5, 20, 6)
    assert np.array_equal(full["xyza"][11:12], row["xyza"])
    assert np.array_equal(full["xyza"][:, 5:6], col["xyza"])
    assert np.array_equal(full["steps"][:, 5:6], col["steps"])


@pytest.mark.parametrize("max_steps", [1, 2, 3, 50])
def test_tiny_step_budgets(ctx, max_steps):
    """`for i in 1..max_steps` (integrator.rs:143): budget 1 integrates nothing, budget n takes at most n-1 steps;
    a ray that runs out of budget has no terminal event and is classed Captured (scene.rs:157-214)."""
    sc = osc.scene_from_config(_config("C2", adaptive=False), config_opts("C2", width=16, height=16, max_steps=max_steps),
                               ROOT)
    dev, ora = both(ctx, sc)
    compare(dev, ora, f"C2 16x16 max_steps={max_steps}", min_fraction=1.0)
    assert dev["steps"].max() <= max(max_steps - 1, 0)
    assert np.array_equal(dev["steps"], ora["steps"])


@pytest.mark.parametrize("name", ["C1", "C2", "C3"])
def test_scene_without_objects(ctx, name):
    """Only the horizon and the celestial sphere: the chord pre-filter has nothing to test."""
    kw = dict(width=20, height=20)
    if name == "C3":
        kw["max_steps"] = 3000
    sc = osc.scene_from_config(_config(name, adaptive=False, drop_objects=True), config_opts(name, **kw), ROOT)
    dev, ora = both(ctx, sc)
    compare(dev, ora, f"{name} without objects", min_fraction=1.0)
    assert not (dev["ray_class"] == abi.CLASS_HIT).any()


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_exact_build_is_tighter_and_fma_build_agrees(ctx, name):
    """grrt_set_option(ctx, "fma", 0) selects the contraction-free kernels for every geometry: they follow the
    oracle's arithmetic operation for operation (only sincos differs from glibc's, by <= 2 ulp, which the geodesic
    amplifies to ~1e-8 at most), so on pixels with the oracle's step count the tolerance drops to 1e-6 (measured:
    C1 3e-11, C2 1.3e-8 at the 99th percentile); the FMA build stays inside the 1e-4 gate."""
    sc = oracle_config_scene(name, width=40, height=40)
    fast, ora = both(ctx, sc)
    ctx.set_option("fma", 0)
    try:
        exact = ctx.render(sc)
    finally:
        ctx.set_option("fma", 1)
    compare(exact, ora, f"{name} exact build")
    compare(fast, ora, f"{name} FMA build")
    same = exact["steps"] == ora["steps"]
    assert same.mean() > 0.99
    denom = np.maximum(np.abs(ora["xyza"][..., :3]).max(axis=-1), 1e-300)
    rel = np.abs(exact["xyza"][..., :3] - ora["xyza"][..., :3]).max(axis=-1) / denom
    assert np.quantile(rel[same], 0.99) < 1e-6


def test_pinned_host_buffers_take_the_overlapped_read_back(ctx):
    """With page-locked output buffers grrt_render copies the base frame out while the supersampling pass runs and then
    patches the rewritten pixels; the frame must equal the one read back into pageable memory."""
    import torch
    for name, kw in (("C2", dict(width=96, height=64)), ("C1", dict(width=64, height=48))):
        sc = oracle_config_scene(name, **kw)
        want = ctx.render(sc)
        assert want["stats"]["pixels_supersampled"] > 0
        n = kw["width"] * kw["height"]
        pinned = {"xyza": torch.zeros((n, 4), dtype=torch.float64).pin_memory().numpy(),
                  "ray_class": torch.zeros(n, dtype=torch.uint8).pin_memory().numpy(),
                  "stop": torch.zeros(n, dtype=torch.uint8).pin_memory().numpy(),
                  "steps": torch.zeros(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
        for _ in range(2):
            got = ctx.render(sc, out=pinned)
            for key in ("xyza", "ray_class", "stop", "steps"):
                assert np.array_equal(got[key], want[key]), (name, key)
        crop = ctx.render(sc, 5, 7, 37, 61)
        m = 32 * 54
        sub = {k: v[:m] if k != "xyza" else v[:m] for k, v in pinned.items()}
        got = ctx.render(sc, 5, 7, 37, 61, out=sub)
        assert np.array_equal(got["xyza"], crop["xyza"]) and np.array_equal(got["steps"], crop["steps"])


def test_pilot_ordered_base_pass_is_bit_identical(ctx):
    """Scenes with --max-steps above 65536 trace their tiles in the order of a 1/8-resolution pilot pass (and deal them
    to the devices in that order); the frame, the step counts and the statistics must not change by a bit."""
    import tomllib
    with open(os.path.join(ROOT, CONFIGS["C3"]["toml"]), "rb") as f:
        cfg = tomllib.load(f)
    cfg["adaptive_sampling"] = {"enabled": False}
    sc = osc.scene_from_config(cfg, config_opts("C3", width=96, height=64, max_steps=70000), ROOT)
    old = os.environ.get("GRRT_PILOT")
    try:
        os.environ["GRRT_PILOT"] = "0"
        plain = ctx.render(sc)
        os.environ["GRRT_PILOT"] = "1"
        piloted = ctx.render(sc)
        crop_p = ctx.render(sc, 3, 5, 61, 93)          # a section large enough for the pilot (58 x 88 pixels)
        os.environ["GRRT_PILOT"] = "0"
        crop = ctx.render(sc, 3, 5, 61, 93)
    finally:
        if old is None:
            os.environ.pop("GRRT_PILOT", None)
        else:
            os.environ["GRRT_PILOT"] = old
    for key in ("xyza", "ray_class", "stop", "steps"):
        assert np.array_equal(piloted[key], plain[key]), key
        assert np.array_equal(crop_p[key], crop[key]), key
    for key in ("accepted_steps", "attempts", "stop_histogram", "rays_base"):
        assert piloted["stats"][key] == plain["stats"][key], key     # the pilot's own rays are not counted
    assert piloted["steps"].max() > 20000                              # some rays outlive the pilot's budget cap
