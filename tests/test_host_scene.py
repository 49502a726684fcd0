"""Host-side scene assembly of the product (C++, libgrrt.so) against the oracle's independent assembly
(Python + oracle library): TOML -> RenderConfig -> camera tetrad, LUTs, textures -> GrrtScene.
Reference: src/cli/shared.rs:131-321, src/rendering/camera.rs:151-196, src/configuration.rs."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from gr_raytracer_b200 import abi, api
from oracle import scene as osc
from helpers import ROOT, CONFIGS, config_opts, oracle_config_scene


def struct_diffs(a, b, path=""):
    out = []
    for name, _ in a._fields_:
        if name.startswith("_pad"):
            continue
        va, vb = getattr(a, name), getattr(b, name)
        if isinstance(va, C.Structure):
            out += struct_diffs(va, vb, path + name + ".")
        elif isinstance(va, C.Array):
            if isinstance(va[0], C.Structure):
                for i in range(len(va)):
                    out += struct_diffs(va[i], vb[i], f"{path}{name}[{i}].")
            elif isinstance(va[0], C.Array):
                for i in range(len(va)):
                    if list(va[i]) != list(vb[i]):
                        out.append(f"{path}{name}[{i}]")
            elif list(va) != list(vb):
                out.append(path + name)
        elif isinstance(va, C._Pointer) or va is None:
            continue
        elif va != vb and not (isinstance(va, float) and va != va and vb != vb):  # NaN == NaN here
            out.append(path + name)
    return out


def buffers_equal(h, o):
    def arr(p, n):
        return np.ctypeslib.as_array(p, (n,))
    for k in range(h.n_objects):
        a, b = h.objects[k], o.objects[k]
        if a.kind != abi.OBJ_SPHERE and a.temperature.kind == abi.TEMP_LUT:
            assert np.array_equal(arr(a.temperature.lut_r, 1000), arr(b.temperature.lut_r, 1000))
            assert np.array_equal(arr(a.temperature.lut_t, 1000), arr(b.temperature.lut_t, 1000))
        if a.texture.kind == abi.TEX_BITMAP:
            n = a.texture.width * a.texture.height * 4
            assert np.array_equal(arr(a.texture.rgba, n), arr(b.texture.rgba, n))
    if h.bb_lut_n:
        assert np.array_equal(arr(h.bb_lut_xyz, 3000), arr(o.bb_lut_xyz, 3000))
        assert np.array_equal(arr(h.bb_lut_log_t, 1000), arr(o.bb_lut_log_t, 1000))
    if h.celestial.kind == abi.TEX_BITMAP:
        n = h.celestial.width * h.celestial.height * 4
        assert np.array_equal(arr(h.celestial.rgba, n), arr(o.celestial.rgba, n))


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_baseline_scenes_assemble_bit_identically(name):
    kw = dict(width=64, height=48)
    hs = api.HostScene(os.path.join(ROOT, CONFIGS[name]["toml"]), config_opts(name, **kw), ROOT)
    os_ = oracle_config_scene(name, **kw)
    assert struct_diffs(hs.c, os_.c) == []
    buffers_equal(hs.c, os_.c)


@pytest.mark.parametrize("toml,kw", [
    ("scenes/euclidean.toml", dict(camera_position=[10.0, 0.0, 0.8])),
    ("scenes/euclidean-spherical.toml", dict(camera_position=[10.0, 0.0, 0.8])),
    ("scenes/kerr-bl.toml", dict(camera_position=[-8.0, 1.0, 2.0], theta=-3.0, psi=0.3, phi=0.2, show_sampling_mask=1)),
])
def test_other_scenes_and_camera_angles(toml, kw):
    opts = abi.default_cli_opts(width=33, height=21, **kw)
    hs = api.HostScene(os.path.join(ROOT, toml), opts, ROOT)
    os_ = osc.scene_from_toml(os.path.join(ROOT, toml), opts, ROOT)
    assert struct_diffs(hs.c, os_.c) == []
    buffers_equal(hs.c, os_.c)


TEXT = """
celestial_temperature = 0.0
camera_velocity = "Zamo"
[celestial_texture.Checker]
beaming_exponent = 0.0
width = 10.0
height = 10.0
color1 = [0, 255, 0]
color2 = [0, 100, 0]
[geometry_type.KerrBL]
radius = 1.0   # r_s
a = 0.3
horizon_epsilon = 1e-4
[adaptive_sampling]
samples_per_axis = 2
minimum_luminance = 1.0
exclude_background_contrast = false
[[objects]]
[objects.Sphere]
radius = 1.5
position = [0.0, 4.0, 1.0]
temperature = 3000.0
[objects.Sphere.texture.BlackBody]
beaming_exponent = 1.0
[[objects]]
[objects.VolumetricDisc]
inner_radius = 3.0
outer_radius = 6.0
temperature = 4000.0
axis = [0.1, 0.0, 1.0]
num_octaves = 3
perlin_seed = 7
max_steps = 100
step_size = 0.01
thickness = 0.1
density_multiplier = 10.0
brightness_reference_temperature = 1000.0
absorption = 0.3
scattering = 0.4
noise_scale = [6.0, 2.0, 3.0]
noise_offset = 0.0
texture = { Checker = { beaming_exponent = 2.0, width = 4.0, height = 4.0, color1 = [255, 0, 0], color2 = [9, 9, 9] } }
"""


def test_reference_style_toml_with_optional_fields():
    """The reference's scene files use [a.b.c] tables; optional fields: camera_velocity, adaptive_sampling, axis, perlin_seed."""
    opts = abi.default_cli_opts(width=20, height=20, camera_position=[-9.0, 0.0, 1.0], theta=-3.14)
    hs = api.HostScene(TEXT, opts, ROOT, is_text=True)
    os_ = osc.scene_from_toml(TEXT, opts, ROOT, is_text=True)
    assert struct_diffs(hs.c, os_.c) == []
    buffers_equal(hs.c, os_.c)
    ad = hs.c.adaptive
    assert (ad.samples_per_axis, ad.has_minimum_luminance, ad.minimum_luminance, ad.exclude_background_contrast) == (2, 1, 1.0, 0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/scene-definitions"), reason="reference mount absent")
def test_reference_scene_files_parse_like_tomllib():
    """Every scene file of the reference goes through the product's TOML reader and yields the scene the oracle
    builds from tomllib's parse (textures resolved against this repo's synthetic resources/)."""
    files = sorted(glob.glob("/root/reference/scene-definitions/*.toml"))
    assert len(files) >= 14
    for f in files:
        opts = abi.default_cli_opts(width=16, height=16, camera_position=[-12.0, 0.0, 1.5])
        try:
            os_ = osc.scene_from_toml(f, opts, ROOT)
        except osc.OracleError as e:
            with pytest.raises(api.GrrtError):
                api.HostScene(f, opts, ROOT)
            continue
        hs = api.HostScene(f, opts, ROOT)
        assert struct_diffs(hs.c, os_.c) == [], f


@pytest.mark.parametrize("patch,code", [
    ("[adaptive_sampling]\nsamples_per_axis = 0\n", abi.ERR_INVALID_CONFIGURATION),
    ("[adaptive_sampling]\nluminance_contrast_threshold = 1.5\n", abi.ERR_INVALID_CONFIGURATION),
    ("[adaptive_sampling]\nminimum_luminance = -1.0\n", abi.ERR_INVALID_CONFIGURATION),
    ("camera_velocity = { Explicit = { components = [2.0, 0.0, 0.0, 0.0] } }\n", abi.ERR_INVALID_CONFIGURATION),
    ("bogus = = 3\n", abi.ERR_TOML),
])
def test_configuration_errors(patch, code):
    """configuration.rs:61-93, cli/shared.rs:48-77."""
    base = open(os.path.join(ROOT, "scenes/kerr-bl.toml")).read()
    text = patch + base if not patch.startswith("[") else base + "\n" + patch
    with pytest.raises(api.GrrtError) as e:
        api.HostScene(text, abi.default_cli_opts(width=8, height=8, camera_position=[-10.0, 0.0, -0.5]), ROOT, is_text=True)
    assert e.value.code == code
    import tomllib
    with pytest.raises((osc.OracleError, tomllib.TOMLDecodeError)):
        osc.scene_from_toml(text, abi.default_cli_opts(width=8, height=8, camera_position=[-10.0, 0.0, -0.5]), ROOT, is_text=True)


def test_missing_files_and_volumetric_validation():
    with pytest.raises(api.GrrtError) as e:
        api.HostScene("/nonexistent/scene.toml", None, ROOT)
    assert e.value.code == abi.ERR_INVALID_CONFIGURATION and "Config file not found" in str(e.value)
    text = open(os.path.join(ROOT, "scenes/schwarzschild.toml")).read().replace("resources/disk.png", "resources/none.png")
    with pytest.raises(api.GrrtError) as e:
        api.HostScene(text, None, ROOT, is_text=True)
    assert e.value.code == abi.ERR_TEXTURE
    vol = open(os.path.join(ROOT, "scenes/schwarzschild-volumetric-dense.toml")).read()
    for bad in ("thickness = 0.1", "thickness = 0.0"), ("step_size = 0.002", "step_size = -1.0"), ("max_steps = 50000", "max_steps = 0"), \
               ("outer_radius = 11.0", "outer_radius = 4.0"):
        with pytest.raises(api.GrrtError) as e:
            api.HostScene(vol.replace(*bad), None, ROOT, is_text=True)
        assert e.value.code == abi.ERR_INVALID_CONFIGURATION and "VolumetricDisc requires" in str(e.value)


def test_png_codec_roundtrip(tmp_path):
    from PIL import Image
    L = api.lib()
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    path = str(tmp_path / "a.png")
    assert L.grrt_host_write_png_rgb8(path.encode(), rgb.ctypes.data_as(abi.c_uint8_p), 53, 37) == 0
    assert np.array_equal(np.array(Image.open(path).convert("RGB")), rgb)      # PIL reads what we wrote
    # we read what PIL wrote, for every colour type / bit depth PIL can produce
    rgba = rng.integers(0, 256, (19, 23, 4), dtype=np.uint8)
    cases = {"rgba": Image.fromarray(rgba, "RGBA"), "rgb": Image.fromarray(rgb, "RGB"),
             "gray": Image.fromarray(rgb[..., 0], "L"), "la": Image.fromarray(rgba[..., :2], "LA"),
             "pal": Image.fromarray(rgb, "RGB").quantize(17), "one": Image.fromarray(rgb[..., 0] > 127),
             "i16": Image.fromarray((rgb[..., 0].astype(np.uint16) * 257), "I;16")}
    for name, img in cases.items():
        p = str(tmp_path / (name + ".png"))
        img.save(p)
        out, w, h = abi.c_uint8_p(), C.c_uint32(), C.c_uint32()
        assert L.grrt_host_read_png_rgba8(p.encode(), C.byref(out), C.byref(w), C.byref(h)) == 0, name
        got = np.ctypeslib.as_array(out, (h.value, w.value, 4)).copy()
        L.grrt_host_free(out)
        want = np.array(Image.open(p).convert("RGBA"))
        if name == "i16":
            want = np.dstack([rgb[..., 0]] * 3 + [np.full_like(rgb[..., 0], 255)])
        assert np.array_equal(got, want), name


@pytest.mark.skipif(not os.path.isdir("/root/reference/resources"), reason="reference mount absent")
def test_decodes_reference_textures_like_pil():
    from PIL import Image
    L = api.lib()
    for f in sorted(glob.glob("/root/reference/resources/*.png")):
        out, w, h = abi.c_uint8_p(), C.c_uint32(), C.c_uint32()
        assert L.grrt_host_read_png_rgba8(f.encode(), C.byref(out), C.byref(w), C.byref(h)) == 0, f
        got = np.ctypeslib.as_array(out, (h.value, w.value, 4)).copy()
        L.grrt_host_free(out)
        assert np.array_equal(got, np.array(Image.open(f).convert("RGBA"))), f


def test_hdr_writer(tmp_path):
    """raytracer.rs:468-480: XYZ written as f32 RGBE."""
    from PIL import Image
    xyza = np.zeros((4, 5, 4))
    xyza[..., 0], xyza[..., 1], xyza[..., 2], xyza[..., 3] = 0.5, 1.25, 300.0, 1.0
    xyza[0, 0, :3] = 0.0
    p = str(tmp_path / "a.hdr")
    assert api.lib().grrt_host_write_hdr_xyz(p.encode(), xyza.ctypes.data_as(abi.c_double_p), 5, 4) == 0
    data = open(p, "rb").read()
    assert data.startswith(b"#?RADIANCE") and b"-Y 4 +X 5" in data
    px = np.frombuffer(data[-4 * 20:], dtype=np.uint8).reshape(4, 5, 4)
    assert tuple(px[0, 0]) == (0, 0, 0, 0)
    r, g, b, e = (int(v) for v in px[1, 1])
    scale = 2.0 ** (e - 128) / 256.0
    assert abs(r * scale - 0.5) <= scale and abs(g * scale - 1.25) <= scale and abs(b * scale - 300.0) <= scale


@pytest.mark.parametrize("name", ["C1", "C2", "C3"])
def test_camera_update_equals_fresh_assembly(name):
    """grrt_host_scene_set_camera (animation frames): recomputing only the camera of an assembled scene gives the
    bytes of a scene assembled from scratch with those options, and a failed update keeps the previous camera."""
    toml = os.path.join(ROOT, CONFIGS[name]["toml"])
    kw = dict(camera_position=(-9.0, 2.0, 0.4), phi=0.05, theta=-3.1, psi=0.02)
    sc = api.HostScene(toml, config_opts(name, width=16, height=12), ROOT)
    sc.set_camera(**kw)
    fresh = api.HostScene(toml, config_opts(name, width=16, height=12, **kw), ROOT)
    assert bytes(sc.c.camera) == bytes(fresh.c.camera)
    assert sc.c.n_objects == fresh.c.n_objects and sc.c.integration.max_steps == fresh.c.integration.max_steps
    if name != "C3":   # Kerr-Schild coordinates have a static observer only outside the ergosphere; r = 0.3 r_s has none
        before = bytes(sc.c.camera)
        with pytest.raises(api.GrrtError):
            sc.set_camera(camera_position=(0.3, 0.0, 0.0))
        assert bytes(sc.c.camera) == before
    sc.close()
    fresh.close()


def test_blackbody_utilities_match_oracle_and_reference_kat(tmp_path):
    """cli/blackbody.rs on the host code of libgrrt: the XYZ integral and xyz_to_srgb agree with the oracle bit for
    bit, reproduce the reference's known colours (black_body_radiation.rs:63-73) and the `blackbody` subcommand
    prints them in the reference's format."""
    import subprocess
    L, O = api.lib(), osc.lib()
    for temp, shift in ((1000.0, 1.0), (10000.0, 1.0), (4500.0, 0.7), (6500.0, 1.8)):
        got, want = (C.c_double * 3)(), (C.c_double * 4)()
        L.grrt_host_blackbody_xyz(temp, shift, got)
        O.orc_blackbody_xyz(temp, shift, want)
        assert list(got) == list(want)[:3]
        for exposure in (1.0, 1.0 / sum(got)):
            a, b = (C.c_uint8 * 3)(), (C.c_uint8 * 3)()
            L.grrt_host_xyz_to_srgb8(got, exposure, a)
            O.orc_xyz_to_srgb(want, exposure, b)
            assert tuple(a) == tuple(b)
    for temp, exp in ((1000.0, (255, 60, 0)), (10000.0, (137, 146, 172))):
        xyz, rgb = (C.c_double * 3)(), (C.c_uint8 * 3)()
        L.grrt_host_blackbody_xyz(temp, 1.0, xyz)
        L.grrt_host_xyz_to_srgb8(xyz, 1.0 / sum(xyz), rgb)
        assert tuple(rgb) == exp
    exe = os.path.join(ROOT, "gr_raytracer_b200", "grrt")
    r = subprocess.run([exe, "blackbody", "--temperature", "6500", "-r", "0.9"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    xyz, rgb = (C.c_double * 3)(), (C.c_uint8 * 3)()
    L.grrt_host_blackbody_xyz(6500.0, 0.9, xyz)
    L.grrt_host_xyz_to_srgb8(xyz, 1.0, rgb)
    assert lines[0] == "Blackbody color at T=6500K (redshift=0.9):"
    assert lines[1] == "XYZ:  %.4f, %.4f, %.4f" % tuple(xyz)
    assert lines[2] == "sRGB: R=%d, G=%d, B=%d" % tuple(rgb)
    assert subprocess.run([exe, "blackbody"], capture_output=True, text=True).returncode != 0
