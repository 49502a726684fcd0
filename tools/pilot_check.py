"""The pilot-ordered base pass must not change a single bit of the frame: renders a Kerr-Schild frame with a step budget
above the pilot threshold with and without the pilot (GRRT_PILOT=0) and compares.

    python tools/pilot_check.py [width height max_steps]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np
    from gr_raytracer_b200 import api
    from helpers import CONFIGS, config_opts
    w, h, ms, out = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    res = {}
    with api.Context(0) as ctx:
        for name in ("C3", "C4"):
            sc = api.HostScene(os.path.join(ROOT, CONFIGS[name]["toml"]), config_opts(name, width=w, height=h, max_steps=ms), ROOT)
            ctx.render(sc, 0, 0, 8, 8)   # warm the kernels
            r = ctx.render(sc)
            st = r["stats"]
            print(f"{name} {w}x{h} max_steps {ms} GRRT_PILOT={os.environ.get('GRRT_PILOT', '1')}: total {st['device_ms_total']:.0f} ms "
                  f"(base {st['device_ms_base']:.0f}, supersample {st['device_ms_supersample']:.0f}), steps {st['accepted_steps']}", flush=True)
            for k in ("xyza", "ray_class", "stop", "steps"):
                res[name + "_" + k] = r[k]
            sc.close()
    np.savez(out, **res)
    sys.exit(0)

import numpy as np
w, h, ms = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 256, 100000)))
files = []
for pilot in ("1", "0"):
    out = f"/tmp/pilot_check_{pilot}.npz"
    files.append(out)
    subprocess.check_call([sys.executable, __file__, "--child", str(w), str(h), str(ms), out], env=dict(os.environ, GRRT_PILOT=pilot))
a, b = np.load(files[0]), np.load(files[1])
same = all(np.array_equal(a[k], b[k]) for k in a.files)
print("frames with and without the pilot pass are bit-identical:", same)
assert same
