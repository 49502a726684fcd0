import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, 'tests'))
from gr_raytracer_b200 import abi, api
from helpers import CONFIGS, config_opts, ROOT
import tomllib

ctx = api.Context(0)
text = open(os.path.join(ROOT, CONFIGS['C3']['toml'])).read() + "\n[adaptive_sampling]\nenabled = false\n"
sc = api.HostScene(text, config_opts('C3', width=224, height=224, max_steps=3000), ROOT, is_text=True)
for i in range(2):
    out = ctx.render(sc)
st = out['stats']
print('C3 224x224 max_steps 3000: base %.1f ms, steps %.3e, steps/s %.3e' % (st['device_ms_base'], st['accepted_steps'], st['accepted_steps']/(st['device_ms_base']*1e-3)))
