"""Builds gr_raytracer_b200/libgrrt.so (and the `grrt` CLI) in-tree with nvcc for sm_100a only.

The ray-tracing translation unit is compiled with -fmad=false: see csrc/grrt_device.cuh.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgrrt.so")
CLI = os.path.join(HERE, "grrt")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# (source, contract a*b+c into FMA?)  see csrc/grrt_kern_exact.cu / grrt_kern_fast.cu
CUDA_SOURCES = ["grrt_api.cu", "grrt_kern_exact.cu", "grrt_kern_fast.cu"]
FMAD = {"grrt_api.cu": "false", "grrt_kern_exact.cu": "false", "grrt_kern_fast.cu": "true"}
HOST_SOURCES = ["host/toml.cpp", "host/png.cpp", "host/scene_build.cpp", "host/host_api.cpp"]
CLI_SOURCES = ["host/cli_main.cpp"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx():
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); nvcc wants the real system compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _all_deps():
    deps = [os.path.join(HERE, "..", "include", "grrt.h")]
    for root, _, files in os.walk(CSRC):
        deps += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp"))]
    return deps


def build(force=False, verbose=False):
    """Compile every CUDA/host source into libgrrt.so; returns the library path."""
    deps = _all_deps()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    common = ["-std=c++17", "-O3", "-lineinfo", "-ccbin", _host_cxx(), "-Xcompiler", "-fPIC", "-I", os.path.join(HERE, "..", "include")]
    objs, jobs = [], []
    for src in CUDA_SOURCES + [s for s in HOST_SOURCES if os.path.exists(os.path.join(CSRC, s))]:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.replace("/", "_") + ".o")
        objs.append(obj)
        if force or _newer(obj, deps):
            cmd = [nvcc] + ARCH + common + ["-c", path, "-o", obj]
            if src.endswith(".cu"):
                cmd += ["-fmad=" + FMAD[src]] + (["-Xptxas", "-v"] if verbose else [])
                if os.environ.get("GRRT_MB"):   # experiment knob: resident CTAs/SM the register allocator targets
                    cmd += ["-DGRRT_MB=" + os.environ["GRRT_MB"]]
                for knob in ("GRRT_MBV", "GRRT_MB_BL"):
                    if os.environ.get(knob):
                        cmd += ["-D%s=%s" % (knob, os.environ[knob])]
            else:
                cmd += ["-Xcompiler", "-ffp-contract=off"]
            jobs.append(cmd)
    if jobs:  # the translation units are independent: compile them concurrently
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            return cmd, r.returncode, r.stdout
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 2)) as ex:
            results = list(ex.map(run, jobs))
        for cmd, rc, out in results:
            if verbose or rc != 0:
                print(out, flush=True)
            if rc != 0:
                raise subprocess.CalledProcessError(rc, cmd)
    if force or _newer(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-ccbin", _host_cxx(), "-o", LIB] + objs + ["-lz", "-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    cli_srcs = [os.path.join(CSRC, s) for s in CLI_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if cli_srcs and (force or _newer(CLI, cli_srcs + [LIB])):
        cmd = [_host_cxx(), "-std=c++17", "-O2", "-I", os.path.join(HERE, "..", "include")] + cli_srcs + \
              ["-o", CLI, "-L", HERE, "-lgrrt", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
