"""The `grrt` command line (the reference's clap surface, src/cli/cli.rs) end to end on the GPU: render to PNG/HDR,
render-ray / render-ray-at CSV export, error exits."""
import os
import subprocess

import numpy as np
import pytest

from gr_raytracer_b200 import abi, api
from oracle import scene as osc
from helpers import ROOT, CONFIGS, config_opts

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "gr_raytracer_b200", "grrt")


def run(args, cwd):
    return subprocess.run([EXE, "--resource-root", ROOT] + args, cwd=cwd, capture_output=True, text=True)


def test_render_png_matches_oracle_tonemap(tmp_path):
    from PIL import Image
    args = ["--width", "40", "--height", "30", "--camera-position=-10,0,-0.5", "--theta=-3.14159",
            "--config-file", os.path.join(ROOT, "scenes/kerr-bl.toml"), "render", "--filename", "out.png"]
    r = run(args, tmp_path)
    assert r.returncode == 0, r.stderr
    img = np.array(Image.open(tmp_path / "out.png"))
    assert img.shape == (30, 40, 3)
    sc = osc.scene_from_toml(os.path.join(ROOT, "scenes/kerr-bl.toml"), config_opts("C2", width=40, height=30), ROOT)
    ora = osc.render(sc)
    ref = np.zeros((30 * 40, 3), dtype=np.uint8)
    osc.lib().orc_tonemap(osc.dptr(np.ascontiguousarray(ora["xyza"].reshape(-1, 4))), 30 * 40, abi.TONEMAP_REINHARD, 1.0, osc.u8ptr(ref))
    diff = np.abs(img.reshape(-1, 3).astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01
    # section render + HDR output
    r = run(args[:-2] + ["--filename", "crop.hdr", "--from-row", "5", "--to-row", "25", "--from-col", "10", "--to-col", "30"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "crop.hdr", "rb").read(10) == b"#?RADIANCE"


def test_render_ray_csv_matches_oracle(tmp_path):
    cfg = os.path.join(ROOT, "scenes/schwarzschild.toml")
    r = run(["--max-steps", "10", "--config-file", cfg, "render-ray", "--row", "250", "--col", "250"], tmp_path)
    assert r.returncode == 0, r.stderr
    lines = open(tmp_path / "rendered-ray.csv").read().split("\n")
    assert lines[0] == "i,t,tau,x,y,z" and len(lines) == 12   # cli/schwarzschild.rs:133-172: header + 10 steps + trailing newline
    sc = osc.scene_from_toml(cfg, abi.default_cli_opts(max_steps=10), ROOT)
    steps, _, _ = osc.trace_ray(sc, pixel=(250, 250))
    for k, line in enumerate(lines[1:11]):
        f = [float(v) for v in line.split(",")]
        cart = (osc.C.c_double * 4)()
        osc.lib().orc_step_to_cartesian(osc.C.byref(sc.c.geometry), osc.darr(steps["x"][k]), cart)
        assert int(f[0]) == k and f[1] == pytest.approx(steps["t"][k], rel=1e-12, abs=1e-15)
        assert f[2:] == pytest.approx(list(cart), rel=1e-9, abs=1e-12)
    r = run(["--max-steps", "50", "--config-file", os.path.join(ROOT, "scenes/kerr.toml"), "render-ray-at", "--position=-10,0,2",
             "--direction=1,0,0", "--filename", "at.csv"], tmp_path)
    assert r.returncode == 0, r.stderr
    # same ray through the oracle: momentum from the unboosted tetrad (cli/kerr.rs:77-100), then Integrator::integrate
    import ctypes as C
    sck = osc.scene_from_toml(os.path.join(ROOT, "scenes/kerr.toml"), abi.default_cli_opts(max_steps=50), ROOT)
    pos, mom, err = (C.c_double * 4)(), (C.c_double * 4)(), C.create_string_buffer(256)
    assert api.lib().grrt_host_ray_at(C.byref(sck.c.geometry), osc.darr([-10, 0, 2]), osc.darr([1, 0, 0]), pos, mom, err, 256) == 0
    o_steps, _, _ = osc.trace_ray(sck, list(pos), list(mom))
    at_lines = open(tmp_path / "at.csv").read().split("\n")
    assert len(at_lines) == len(o_steps) + 2
    last = [float(v) for v in at_lines[len(o_steps)].split(",")]
    assert last[3:] == pytest.approx(list(o_steps["x"][-1][1:]), rel=1e-12)   # Kerr-Schild: bit-identical trajectory


def test_cli_errors(tmp_path):
    assert run(["render"], tmp_path).returncode != 0                                   # no config file
    r = run(["--config-file", "missing.toml", "render"], tmp_path)
    assert r.returncode != 0 and "Config file not found" in r.stderr                   # main.rs:67-73
    r = run(["--camera-position=1,2", "--config-file", os.path.join(ROOT, "scenes/kerr.toml"), "render"], tmp_path)
    assert r.returncode != 0 and "length 3" in r.stderr                                # main.rs:85-89


def test_blackbody_spectrum_png_matches_oracle(tmp_path):
    """cli/blackbody.rs run_blackbody_spectrum: temperature along x, redshift along y, tone-mapped on the device."""
    from PIL import Image
    W, H = 24, 10
    r = run(["--tone-mapping", "global-linear", "blackbody-spectrum", "--width", str(W), "--height", str(H), "--min-temperature",
             "1500", "--max-temperature", "9000", "--filename", "bb.png"], tmp_path)
    assert r.returncode == 0, r.stderr
    img = np.array(Image.open(tmp_path / "bb.png"))
    assert img.shape == (H, W, 3)
    xyza = np.zeros((H * W, 4))
    out = (osc.C.c_double * 4)()
    for i in range(H * W):
        t = 1500.0 + (i % W) * (9000.0 - 1500.0) / (W - 1.0)
        g = 0.5 + (i // W) * (2.0 - 0.5) / (H - 1.0)
        osc.lib().orc_blackbody_xyz(t, g, out)
        xyza[i] = [out[0], out[1], out[2], 1.0]
    ref = np.zeros((H * W, 3), dtype=np.uint8)
    osc.lib().orc_tonemap(osc.dptr(xyza), H * W, abi.TONEMAP_GLOBAL_LINEAR, 1.0, osc.u8ptr(ref))
    diff = np.abs(img.reshape(-1, 3).astype(int) - ref.astype(int))
    assert diff.max() <= 1


def test_plain_c_consumer(tmp_path):
    """tests/c_abi_smoke.c -- C99, nothing but include/grrt.h and libgrrt.so -- renders a frame through
    grrt_host_scene_from_toml_file -> grrt_render -> grrt_tonemap; its numbers equal the Python binding's."""
    import torch
    exe = str(tmp_path / "c_abi_smoke")
    lib_dir = os.path.join(ROOT, "gr_raytracer_b200")
    cc = subprocess.run(["cc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", lib_dir, "-lgrrt", f"-Wl,-rpath,{lib_dir}"],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    toml = os.path.join(ROOT, "scenes", "kerr-bl.toml")
    want = None
    for n_dev in sorted({1, torch.cuda.device_count()}):
        r = subprocess.run([exe, toml, ROOT, "40", "30", str(n_dev)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        tok = r.stdout.split()
        assert tok[0] == "ok" and (int(tok[1]), int(tok[2])) == (40, 30)
        if want is None:
            from gr_raytracer_b200 import abi, api
            scene = api.HostScene(toml, abi.default_cli_opts(width=40, height=30, camera_position=[-10.0, 0.0, -0.5], theta=-3.14159), ROOT)
            with api.Context(0) as ctx:
                out = ctx.render(scene)
                rgb = ctx.tonemap(out["xyza"])
            st = out["stats"]
            want = (st["rays_base"] + st["rays_supersample"], st["accepted_steps"], float(out["xyza"][..., :3].sum(axis=-1).ravel().sum()),
                    tuple(int(v) for v in rgb[15, 20]))
        assert int(tok[3]) == want[0] and int(tok[4]) == want[1]
        assert abs(float(tok[5]) - want[2]) <= 1e-9 * abs(want[2])      # same pixels, summed in a different order
        assert tuple(int(v) for v in tok[6:9]) == want[3]
