import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from gr_raytracer_b200 import abi, api
from gr_raytracer_b200.renderer import DeviceEngine, FrameRenderer
ROOT='/root/repo'
opts = abi.default_cli_opts(width=1500, height=1500, camera_position=[-10.0, 0.0, -0.5], theta=-3.14159)
scene = api.HostScene(os.path.join(ROOT, 'scenes/kerr-bl.toml'), opts, ROOT)
ctx = api.Context(0)
eng = DeviceEngine(ctx, scene)
fr = FrameRenderer(eng, scene.c.adaptive, 1500, 1500)
E = lambda: torch.cuda.Event(enable_timing=True)
for i in range(4):
    ev = [E() for _ in range(4)]
    torch.cuda.synchronize(); t0=time.perf_counter()
    ev[0].record()
    eng.trace(None, fr.n, fr.rect, 0, fr.xyza, fr.cls, fr.stop, fr.steps); ta=time.perf_counter()
    ev[1].record()
    eng.select(fr.xyza, fr.cls, 1500, 1500, fr.selected, fr.count); tb=time.perf_counter(); n=int(fr.count.item()); tc=time.perf_counter()
    ev[2].record()
    eng.trace(fr.selected, n, fr.rect, 4, fr.xyza); td=time.perf_counter()
    ev[3].record()
    torch.cuda.synchronize(); te=time.perf_counter()
    print('events: base %.1f select %.1f ss %.1f total %.1f | host: launch %.1f select %.1f item %.1f ss %.1f sync %.1f wall %.1f' % (
        ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3]),
        (ta-t0)*1e3, (tb-ta)*1e3, (tc-tb)*1e3, (td-tc)*1e3, (te-td)*1e3, (te-t0)*1e3))
for i in range(3):
    t0=time.perf_counter(); out = ctx.render(scene); t1=time.perf_counter()
    st = out['stats']
    print('grrt_render wall %.1f ms; base %.1f mask %.1f ss %.1f total %.1f' % ((t1-t0)*1e3, st['device_ms_base'], st['device_ms_mask'], st['device_ms_supersample'], st['device_ms_total']))
