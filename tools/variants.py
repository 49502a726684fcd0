"""Kernel experiments: builds variants of the trace kernels (extra -D flags on the two kernel translation units)
into gr_raytracer_b200/variants/libgrrt_<name>.so and times them on the GPU.

    python tools/variants.py build name1="-DFOO -DBAR=2" name2=...      (here, no GPU needed)
    python tools/variants.py run C2:1500:1500 C1:500:500 -- name1 name2  (on the GPU box; 'base' = libgrrt.so)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gr_raytracer_b200")
VAR = os.path.join(PKG, "variants")
sys.path.insert(0, ROOT)


def build(name, defs, only_fast=True):
    from gr_raytracer_b200 import build as b
    os.makedirs(VAR, exist_ok=True)
    objdir = os.path.join(VAR, "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    nvcc = b._nvcc()
    common = ["-std=c++17", "-O3", "-lineinfo", "-ccbin", b._host_cxx(), "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")]
    objs = []
    procs = []
    for src in b.CUDA_SOURCES:
        base_obj = os.path.join(PKG, "build", src + ".o")
        rebuild = src == "grrt_kern_fast.cu" or (not only_fast and src == "grrt_kern_exact.cu")
        if not rebuild:
            objs.append(base_obj)
            continue
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        cmd = [nvcc] + b.ARCH + common + ["-c", os.path.join(b.CSRC, src), "-o", obj, "-fmad=" + b.FMAD[src], "-Xptxas", "-v"] + defs.split()
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out = p.communicate()[0]
        if p.returncode:
            print(out)
            raise SystemExit(f"{name}: {src} failed")
        lines = out.split("\n")
        for i, l in enumerate(lines):
            if "Compiling entry function" in l and "trace_kernel" in l and ("ELb0E" in l):
                k = l.split("'")[1]
                print(f"  [{name}] {k[:60]}: {lines[i + 2].strip()} | {lines[i + 3].strip()}")
    for s in b.HOST_SOURCES:
        objs.append(os.path.join(PKG, "build", s.replace("/", "_") + ".o"))
    lib = os.path.join(VAR, f"libgrrt_{name}.so")
    subprocess.check_call([nvcc] + b.ARCH + ["-shared", "-ccbin", b._host_cxx(), "-o", lib] + objs + ["-lz", "-cudart", "static"])
    return lib


RUNNER = r"""
import sys, os, json, time
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))
from gr_raytracer_b200 import abi, api
from helpers import CONFIGS, config_opts, ROOT
ctx = api.Context(0)
res = {}
for spec in sys.argv[1:]:
    parts = spec.split(':'); name, w, h = parts[0], int(parts[1]), int(parts[2])
    kw = dict(width=w, height=h)
    if len(parts) > 3: kw['max_steps'] = int(parts[3])
    sc = api.HostScene(os.path.join(ROOT, CONFIGS[name]['toml']), config_opts(name, **kw), ROOT)
    best = None
    for rep in range(3):
        out = ctx.render(sc)
        st = out['stats']
        if best is None or st['device_ms_total'] < best['device_ms_total']: best = st
    res[spec] = dict(ms=best['device_ms_total'], base=best['device_ms_base'], ss=best['device_ms_supersample'],
                     attempts=best['attempts'], steps=best['accepted_steps'], samples=best['raymarch_samples'],
                     rays=best['rays_base'] + best['rays_supersample'], errors=best['ray_errors'],
                     csum=float(out['xyza'][..., :3].sum()), hist=list(best['stop_histogram']))
print('RESULT ' + json.dumps(res))
"""


def run(specs, names):
    table = {}
    for n in names:
        env = dict(os.environ)
        if n != "base":
            env["GRRT_LIB"] = os.path.join(VAR, f"libgrrt_{n}.so")
        p = subprocess.run([sys.executable, "-c", RUNNER % (ROOT, ROOT)] + specs, env=env, capture_output=True, text=True, timeout=900)
        line = [l for l in p.stdout.split("\n") if l.startswith("RESULT ")]
        if not line:
            print(n, "FAILED", p.stdout[-500:], p.stderr[-1500:])
            continue
        table[n] = json.loads(line[0][7:])
        for spec, r in table[n].items():
            print(f"{n:24s} {spec:16s} total {r['ms']:8.2f} ms  base {r['base']:8.2f}  ss {r['ss']:7.2f}  attempts/s {r['attempts'] / r['ms'] * 1e3:.4e} "
                  f"steps {r['steps']} att {r['attempts']} err {r['errors']} csum {r['csum']:.10e}", flush=True)
    return table


if __name__ == "__main__":
    if sys.argv[1] == "build":
        for a in sys.argv[2:]:
            name, _, defs = a.partition("=")
            only_fast = True
            if name.startswith("x:"):   # x:name rebuilds the exact translation unit too
                name, only_fast = name[2:], False
            print(build(name, defs, only_fast))
    else:
        i = sys.argv.index("--")
        t = run(sys.argv[2:i], sys.argv[i + 1:])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "variants.json"), "a") as f:
            f.write(json.dumps(t) + "\n")
