"""The `grrt` command line without a GPU: argument validation happens before any device is touched, and a render
fails loudly -- there is no CPU fallback (the reference's flags: src/cli/cli.rs:4-113)."""
import os
import subprocess

import pytest

from helpers import ROOT

EXE = os.path.join(ROOT, "gr_raytracer_b200", "grrt")
COMMON = [EXE, "--resource-root", ROOT, "--width", "8", "--height", "8", "--config-file", os.path.join(ROOT, "scenes/kerr-bl.toml")]


def run(*args):
    return subprocess.run(list(args), capture_output=True, text=True, cwd=ROOT)


@pytest.mark.parametrize("pattern", ["%s.png", "frame_%d_%d.png", "%n", "plain.png", "f_%05x.png", "f_%0d.png"])
def test_output_pattern_is_validated_not_passed_to_printf(pattern, tmp_path):
    pos = tmp_path / "p.csv"
    pos.write_text("x,y,z\n-10,0,-0.5\n")
    r = run(*COMMON, "render-frames", "--camera-positions", str(pos), "--output-pattern", pattern)
    assert r.returncode == 1 and "--output-pattern" in r.stderr and "%d" in r.stderr


@pytest.mark.parametrize("spec", ["0,0x", "3-1", "a", "0,,1", "-1"])
def test_devices_list_is_validated(spec):
    r = run(*COMMON, "--devices", spec, "render")
    assert r.returncode == 1 and "--devices" in r.stderr


def test_unknown_flag_and_missing_subcommand():
    r = run(*COMMON, "--no-such-flag", "render")
    assert r.returncode == 1 and "unexpected argument" in r.stderr
    r = run(*COMMON)
    assert r.returncode == 1 and "subcommand is required" in r.stderr


def test_render_without_a_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for extra in (["--device", "0"], ["--devices", "0-1"]):
        r = run(*COMMON, *extra, "render", "--filename", "/tmp/never_written.png")
        assert r.returncode == 1 and ("no usable CUDA device" in r.stderr or "--devices" in r.stderr)
        assert not os.path.exists("/tmp/never_written.png")
