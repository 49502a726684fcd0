"""torchrun --nproc-per-node N tools/mgpu_phases.py: per-rank phase times of the sharded C2 frame."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import torch, torch.distributed as dist
from gr_raytracer_b200 import api
from gr_raytracer_b200.renderer import DeviceEngine, FrameRenderer
from helpers import CONFIGS, config_opts, ROOT
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = api.Context(local)
sc = api.HostScene(os.path.join(ROOT, CONFIGS["C2"]['toml']), config_opts("C2"), ROOT)
eng = DeviceEngine(ctx, sc)
fr = FrameRenderer(eng, sc.c.adaptive, 1500, 1500, rank, world)
for _ in range(3):
    fr.render()
torch.cuda.synchronize(); dist.barrier()
fr.profile = True
acc = {}
for _ in range(3):
    dist.barrier()
    fr.render()
    for k, v in fr.phase_ms.items():
        acc[k] = acc.get(k, 0.0) + v / 3
keys = list(acc)
t = torch.tensor([acc[k] for k in keys], dtype=torch.float64, device="cuda")
allt = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allt, t)
if rank == 0:
    print("phase", " ".join(f"{k:>12s}" for k in keys))
    for r in range(world):
        print(f"rank{r}", " ".join(f"{v:12.2f}" for v in allt[r].tolist()))
if rank == 0:
    print('peer-store path' if fr.p2p else 'all-gather path')
fr.close()
dist.barrier(); dist.destroy_process_group()
