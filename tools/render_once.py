"""One frame of a BASELINE.json configuration through the C ABI (the command profiled under ncu).

    python tools/render_once.py C2 [width height [max_steps]]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gr_raytracer_b200 import api  # noqa: E402
from helpers import CONFIGS, config_opts  # noqa: E402

name = sys.argv[1]
kw = {}
if len(sys.argv) > 3:
    kw = dict(width=int(sys.argv[2]), height=int(sys.argv[3]))
if len(sys.argv) > 4:
    kw["max_steps"] = int(sys.argv[4])
scene = api.HostScene(os.path.join(ROOT, CONFIGS[name]["toml"]), config_opts(name, **kw), ROOT)
with api.Context(0) as ctx:
    out = ctx.render(scene)
st = out["stats"]
print(f"{name}: {st['device_ms_total']:.2f} ms (base {st['device_ms_base']:.2f}, supersample {st['device_ms_supersample']:.2f}), "
      f"rays {st['rays_base']}+{st['rays_supersample']}, steps {st['accepted_steps']}, attempts {st['attempts']}")
