"""Two (or more) GPUs in one box: the sharded frame -- peer-store path and NCCL all-gather path -- must equal the
single-GPU frame bit for bit.  Skipped on single-GPU boxes; the host-side sharding logic is covered
on CPU by tests/test_distributed_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_rank_frame_is_bit_identical(p2p):
    env = dict(os.environ, GRRT_P2P=p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571" if p2p == "1" else "29572", os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("bit-identical to 1-GPU frame: True") == 2
    assert ("peer-store path" in r.stdout) == (p2p == "1")
    assert r.stdout.count("is intact and B is right: True") == 3     # back-to-back frames do not race


def _frames_equal(a, b):
    import numpy as np
    return all(np.array_equal(a[k], b[k]) for k in ("xyza", "ray_class", "stop", "steps"))


def test_multi_device_context_through_the_c_abi():
    """grrt_init_multi (one blocking, internally multi-GPU grrt_render): the frame of a context over all
    GPUs of the box equals the single-device frame bit for bit -- stats included.  On a single-GPU box the one-device
    group is checked (same code path as grrt_init) and the argument validation."""
    import numpy as np
    from gr_raytracer_b200 import api
    from helpers import CONFIGS, config_opts
    n_gpu = torch.cuda.device_count()
    with api.Context(0) as single:
        for name, w, h in (("C2", 301, 203), ("C1", 96, 64), ("C5", 40, 32)):
            sc = api.HostScene(os.path.join(ROOT, CONFIGS[name]["toml"]), config_opts(name, width=w, height=h), ROOT)
            want = single.render(sc)
            crop_want = single.render(sc, 8, 16, 40, 90) if name == "C2" else None
            for n in sorted({1, min(2, n_gpu), n_gpu}):
                with api.Context(list(range(n))) as group:
                    assert group.device_count == n
                    got = group.render(sc)
                    assert _frames_equal(got, want), (name, n)
                    for key in ("rays_base", "rays_supersample", "pixels_supersampled", "accepted_steps", "attempts",
                                "raymarch_samples", "stop_histogram", "ray_errors", "minimum_luminance"):
                        assert got["stats"][key] == want["stats"][key], (name, n, key)
                    if crop_want is not None:
                        assert _frames_equal(group.render(sc, 8, 16, 40, 90), crop_want)
                    again = group.render(sc)                      # back-to-back frames on the same group
                    assert _frames_equal(again, want)
            sc.close()
    with pytest.raises(api.GrrtError):
        api.Context([0, 0])            # devices must be distinct
    with pytest.raises(api.GrrtError):
        api.Context([])
    print(f"[multi] grrt_init_multi frames bit-identical to the single-device frame on 1..{n_gpu} devices")


def test_cli_devices_flag(tmp_path):
    """`grrt --devices 0-(N-1) render` writes the same PNG as `grrt --device 0 render`."""
    n_gpu = torch.cuda.device_count()
    exe = os.path.join(ROOT, "gr_raytracer_b200", "grrt")
    common = [exe, "--resource-root", ROOT, "--width", "64", "--height", "48", "--camera-position=-10,0,-0.5", "--theta=-3.14159",
              "--config-file", os.path.join(ROOT, "scenes/kerr-bl.toml")]
    a = subprocess.run(common + ["--device", "0", "render", "--filename", "one.png"], cwd=tmp_path, capture_output=True, text=True)
    b = subprocess.run(common + ["--devices", f"0-{n_gpu - 1}", "render", "--filename", "all.png"], cwd=tmp_path, capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert (tmp_path / "one.png").read_bytes() == (tmp_path / "all.png").read_bytes()
    bad = subprocess.run(common + ["--devices", "0,0", "render"], cwd=tmp_path, capture_output=True, text=True)
    assert bad.returncode != 0 and "--devices" in bad.stderr
