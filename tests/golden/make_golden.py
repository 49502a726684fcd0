"""Regenerates the golden fixtures of tests/golden/ with the CPU oracle (run from the repository root:
`python tests/golden/make_golden.py`).  One .npz per BASELINE.json configuration at a size the oracle renders in
seconds: per-pixel linear XYZA, termination reason, ray class, accepted-step count, plus the first steps of one
trajectory.  The oracle itself is pinned to the reference's known answers by tests/test_oracle_kat.py; these files
pin the oracle's *frames* so that a change in either the oracle or the device path shows up as a diff."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import oracle_config_scene  # noqa: E402
from oracle import scene as osc  # noqa: E402

CASES = {
    "C1": dict(width=32, height=24),
    "C2": dict(width=40, height=30),
    "C3": dict(width=16, height=12, max_steps=20000),
    "C4": dict(width=12, height=10, max_steps=20000),
    "C5": dict(width=20, height=16),
}
RAY_PIXEL = {"C1": (10, 14), "C2": (13, 22), "C3": (5, 9), "C4": (4, 6), "C5": (7, 11)}
RAY_STEPS = 400


def render_case(name):
    sc = oracle_config_scene(name, **CASES[name])
    frame = osc.render(sc)
    steps, stop, _ = osc.trace_ray(sc, pixel=RAY_PIXEL[name])
    n = min(RAY_STEPS, len(steps["x"]))
    return dict(xyza=frame["xyza"], stop=frame["stop"], ray_class=frame["ray_class"], steps=frame["steps"],
                pixels_supersampled=np.int64(frame["stats"]["pixels_supersampled"]),
                ray_x=np.asarray(steps["x"][:n]), ray_p=np.asarray(steps["p"][:n]), ray_len=np.int64(len(steps["x"])),
                ray_stop=np.int64(stop))


if __name__ == "__main__":
    for name in CASES:
        data = render_case(name)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **data)
        print(name, data["xyza"].shape, "supersampled", int(data["pixels_supersampled"]), "ray steps", int(data["ray_len"]))
