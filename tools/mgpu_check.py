"""torchrun --nproc-per-node N tools/mgpu_check.py: the N-rank frame must equal the single-GPU grrt_render frame bit for bit."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch, torch.distributed as dist
from gr_raytracer_b200 import api
from gr_raytracer_b200.renderer import DeviceEngine, FrameRenderer
from helpers import CONFIGS, config_opts, ROOT
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = api.Context(local)
for name, w, h in (("C2", 301, 203), ("C1", 96, 64)):
    sc = api.HostScene(os.path.join(ROOT, CONFIGS[name]['toml']), config_opts(name, width=w, height=h), ROOT)
    eng = DeviceEngine(ctx, sc)
    fr = FrameRenderer(eng, sc.c.adaptive, w, h, rank, world)
    c = fr.render()
    torch.cuda.synchronize()
    if rank == 0:
        ref = ctx.render(sc)
        got = fr.xyza.cpu().numpy().reshape(h, w, 4)
        same = np.array_equal(got, ref["xyza"])
        print(f"{name} {w}x{h} world={world}: bit-identical to 1-GPU frame: {same}; supersampled {c['pixels_supersampled']} vs {ref['stats']['pixels_supersampled']}", flush=True)
        assert same
    if rank == 0:
        print('  peer-store path' if fr.p2p else '  all-gather path', flush=True)
    fr.close()
    eng.close()
dist.barrier()
dist.destroy_process_group()
