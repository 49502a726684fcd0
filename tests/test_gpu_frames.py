"""Animation frames: a scene kept resident on the GPU with only the camera replaced per
frame must give, bit for bit, the frame of a scene assembled from scratch for that camera -- through the C ABI
(grrt_host_scene_set_camera / grrt_scene_set_camera / grrt_render_resident) and through `grrt render-frames`."""
import os
import subprocess

import numpy as np
import pytest

from gr_raytracer_b200 import api
from oracle import scene as osc
from helpers import ROOT, CONFIGS, config_opts, rel_err_xyz

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "gr_raytracer_b200", "grrt")
CAMERAS = [((-10.0, 0.0, -0.5), None), ((-9.0, 2.0, 0.4), None), ((-7.5, 4.0, 1.5), (0.05, -3.1, 0.02))]


@pytest.mark.parametrize("name", ["C2", "C1"])
def test_resident_scene_with_replaced_camera_equals_fresh_scene(name):
    ctx = api.Context(0)
    toml = os.path.join(ROOT, CONFIGS[name]["toml"])
    W, H = 48, 36
    resident = api.HostScene(toml, config_opts(name, width=W, height=H), ROOT)
    ds = ctx.upload(resident)
    try:
        for pos, angles in CAMERAS:
            kw = dict(camera_position=pos)
            if angles:
                kw.update(phi=angles[0], theta=angles[1], psi=angles[2])
            resident.set_camera(**kw)
            ctx.set_camera(ds, resident.c.camera)
            got = ctx.render_resident(ds, H, W)
            fresh_opts = config_opts(name, width=W, height=H, **kw)
            fresh = api.HostScene(toml, fresh_opts, ROOT)
            assert bytes(resident.c.camera) == bytes(fresh.c.camera)   # same Camera::new output
            want = ctx.render(fresh)
            for key in ("xyza", "ray_class", "stop", "steps"):
                assert np.array_equal(got[key], want[key]), (name, pos, key)
            assert got["stats"]["pixels_supersampled"] == want["stats"]["pixels_supersampled"]
            fresh.close()
            # ... and the moved camera against the ORACLE's frame for that camera (not only device against device)
            ora = osc.render(osc.scene_from_toml(toml, fresh_opts, ROOT))
            rel = rel_err_xyz(got["xyza"], ora["xyza"])
            black = np.max(np.abs(ora["xyza"][..., :3]), axis=-1) == 0.0
            rel = np.where(black, np.max(np.abs(got["xyza"][..., :3]), axis=-1), rel)
            good = (rel <= 1e-4) & (got["stop"] == ora["stop"]) & (got["ray_class"] == ora["ray_class"])
            print(f"[frames] {name} camera {pos} {angles}: {100 * good.mean():.2f} % of the resident-scene frame within 1e-4 of the oracle")
            assert good.mean() >= 0.999, (name, pos)
            assert (got["steps"] != ora["steps"]).mean() <= 0.01   # the FMA build may flip an accept / reject on a few rays
            assert got["stats"]["pixels_supersampled"] == ora["stats"]["pixels_supersampled"]
    finally:
        ctx.free_scene(ds)
        resident.close()
        ctx.close()


def test_camera_errors_leave_the_scene_usable():
    """A camera inside the horizon has no static observer (cli/shared.rs:68-86): the update fails, the previous
    camera stays."""
    toml = os.path.join(ROOT, CONFIGS["C1"]["toml"])
    sc = api.HostScene(toml, config_opts("C1", width=8, height=8), ROOT)
    before = bytes(sc.c.camera)
    with pytest.raises(api.GrrtError):
        sc.set_camera(camera_position=(0.3, 0.0, 0.0))
    assert bytes(sc.c.camera) == before
    sc.close()


def test_render_frames_cli_matches_single_renders(tmp_path):
    from PIL import Image
    pos = tmp_path / "camera_positions.csv"
    pos.write_text("x,y,z\n-10,0,-0.5\n-9.5,1.0,0.0\n-9,2,0.4,0.0,-3.14159,0.0\n-8,3,1\n")
    common = ["--resource-root", ROOT, "--width", "40", "--height", "30", "--theta=-3.14159",
              "--config-file", os.path.join(ROOT, "scenes/kerr-bl.toml")]
    r = subprocess.run([EXE] + common + ["render-frames", "--camera-positions", str(pos), "--output-pattern", "f_%03d.png"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for k, p in enumerate(["-10,0,-0.5", "-9.5,1.0,0.0", "-9,2,0.4", "-8,3,1"]):
        one = subprocess.run([EXE] + common + [f"--camera-position={p}", "render", "--filename", f"single_{k}.png"],
                             cwd=tmp_path, capture_output=True, text=True)
        assert one.returncode == 0, one.stderr
        a = np.array(Image.open(tmp_path / f"f_{k:03d}.png"))
        b = np.array(Image.open(tmp_path / f"single_{k}.png"))
        assert a.shape == (30, 40, 3) and np.array_equal(a, b), k
    # sharding a movie over processes: --first / --stride / --count
    r = subprocess.run([EXE] + common + ["render-frames", "--camera-positions", str(pos), "--output-pattern", "s_%03d.png",
                                         "--first", "1", "--stride", "2"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    made = sorted(f for f in os.listdir(tmp_path) if f.startswith("s_"))
    assert made == ["s_001.png", "s_003.png"]
    assert np.array_equal(np.array(Image.open(tmp_path / "s_003.png")), np.array(Image.open(tmp_path / "f_003.png")))
