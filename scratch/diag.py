import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from gr_raytracer_b200 import abi, api
from gr_raytracer_b200.renderer import DeviceEngine, FrameRenderer
ROOT='/root/repo'
opts = abi.default_cli_opts(width=1500, height=1500, camera_position=[-10.0, 0.0, -0.5], theta=-3.14159)
scene = api.HostScene(os.path.join(ROOT, 'scenes/kerr-bl.toml'), opts, ROOT)
ctx = api.Context(0)
for i in range(2):
    t0=time.perf_counter(); out = ctx.render(scene); t1=time.perf_counter()
    st = out['stats']
    print('grrt_render wall %.1f ms; base %.1f mask %.1f ss %.1f total %.1f ; att %d' % ((t1-t0)*1e3, st['device_ms_base'], st['device_ms_mask'], st['device_ms_supersample'], st['device_ms_total'], st['attempts']))
eng = DeviceEngine(ctx, scene)
fr = FrameRenderer(eng, scene.c.adaptive, 1500, 1500)
for i in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    eng.trace(None, fr.n, fr.rect, 0, fr.xyza, fr.cls, fr.stop, fr.steps); torch.cuda.synchronize(); t1=time.perf_counter()
    eng.select(fr.xyza, fr.cls, 1500, 1500, fr.selected, fr.count); n=int(fr.count.item()); torch.cuda.synchronize(); t2=time.perf_counter()
    eng.trace(fr.selected, n, fr.rect, 4, fr.xyza); torch.cuda.synchronize(); t3=time.perf_counter()
    print('base %.1f ms, select %.1f ms, ss %.1f ms (n_sel %d)' % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, n))
