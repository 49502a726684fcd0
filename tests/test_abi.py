"""The C-ABI library loads without a GPU, exports every symbol include/grrt.h declares, and the ctypes
mirror (gr_raytracer_b200/abi.py) has the C layout.  No compute entry point is called here."""
import ctypes as C
import os
import re

from gr_raytracer_b200 import abi, api
from helpers import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "grrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grrt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"libgrrt.so lacks {missing}"
    assert b"sm_100a" in L.grrt_version()


def test_ctypes_layout_matches_c():
    L = api.lib()
    L.grrt_abi_sizeof.restype = C.c_size_t
    L.grrt_abi_sizeof.argtypes = [C.c_int]
    structs = [abi.GrrtGeometry, abi.GrrtCamera, abi.GrrtIntegration, abi.GrrtTexture, abi.GrrtTemperature,
               abi.GrrtObject, abi.GrrtAdaptive, abi.GrrtScene, abi.GrrtStep, abi.GrrtStats, abi.GrrtCliOpts]
    for i, s in enumerate(structs):
        assert L.grrt_abi_sizeof(i) == C.sizeof(s), s.__name__


def test_default_cli_opts_match_reference_defaults():
    """src/cli/cli.rs:6-47."""
    o = abi.GrrtCliOpts()
    api.lib().grrt_cli_opts_default(C.byref(o))
    p = abi.default_cli_opts()
    for name, _ in abi.GrrtCliOpts._fields_:
        a, b = getattr(o, name), getattr(p, name)
        assert (list(a) == list(b)) if hasattr(a, "__len__") else (a == b), name
    assert (o.width, o.height, o.max_steps, o.max_radius, o.step_size, o.epsilon) == (500, 500, 20000, 15000.0, 0.01, 0.00001)
    assert list(o.camera_position) == [18.0, 0.0, 0.8] and list(o.sampling_mask_color) == [255, 0, 255]


def test_init_fails_loudly_without_a_device():
    """No CPU fallback: without a CUDA device grrt_init returns GRRT_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    assert api.lib().grrt_init(0, C.byref(h)) == abi.ERR_CUDA
    try:
        api.Context(0)
    except api.GrrtError as e:
        assert e.code == abi.ERR_CUDA
    else:
        raise AssertionError("Context() must raise without a GPU")


def test_header_is_plain_c99_and_a_c_consumer_links(tmp_path):
    """include/grrt.h compiles as pedantic C99 and tests/c_abi_smoke.c links against libgrrt.so alone; without a GPU the
    program assembles the scene on the host and then fails loudly at grrt_init (exit code 3): no CPU fallback."""
    import subprocess
    import torch
    from helpers import ROOT
    exe = str(tmp_path / "c_abi_smoke")
    lib_dir = os.path.join(ROOT, "gr_raytracer_b200")
    cc = subprocess.run(["cc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", lib_dir, "-lgrrt", f"-Wl,-rpath,{lib_dir}"],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    if torch.cuda.is_available():
        return   # the GPU run of the program is tests/test_gpu_cli.py::test_plain_c_consumer
    r = subprocess.run([exe, os.path.join(ROOT, "scenes", "kerr-bl.toml"), ROOT, "16", "12", "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr
