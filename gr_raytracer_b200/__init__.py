"""gr_raytracer_b200 -- B200 (sm_100a) implementation of gr_raytracer's per-pixel geodesic-integration path.

csrc/            CUDA kernels, the C ABI (include/grrt.h) and the C++ host side (TOML, PNG, scene assembly, CLI)
abi.py, api.py   ctypes mirror of the C ABI
renderer.py      frame driver on device-resident buffers, tile-sharded over torch.distributed ranks
build.py         in-tree nvcc build (sm_100a only)
"""
__all__ = ["abi", "api", "renderer", "build"]
