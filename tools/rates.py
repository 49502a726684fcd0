import sys, os, time
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
from gr_raytracer_b200 import abi, api
from helpers import CONFIGS, config_opts, ROOT
ctx = api.Context(0)
specs = sys.argv[1:]
for spec in specs:
    parts = spec.split(':'); name, w, h = parts[:3]
    kw = dict(width=int(w), height=int(h))
    if len(parts) > 3: kw['max_steps'] = int(parts[3])
    sc = api.HostScene(os.path.join(ROOT, CONFIGS[name]['toml']), config_opts(name, **kw), ROOT)
    t0 = time.perf_counter(); out = ctx.render(sc); dt = time.perf_counter() - t0
    st = out['stats']
    print(f"{name} {w}x{h}: wall {dt:.2f}s base {st['device_ms_base']:.0f} ms ss {st['device_ms_supersample']:.0f} ms | rays {st['rays_base']}+{st['rays_supersample']} steps {st['accepted_steps']:.3e} attempts {st['attempts']:.3e} samples {st['raymarch_samples']:.3e} | steps/s {st['accepted_steps']/(st['device_ms_total']*1e-3):.3e} samples/s {st['raymarch_samples']/(st['device_ms_total']*1e-3):.3e} errors {st['ray_errors']} hist {st['stop_histogram']}", flush=True)
