"""The N>1 host logic (tile sharding, base-buffer all-gather, identical mask on every rank, round-robin
supersampling, result all-gather) under `gloo` with world size 2 on the CPU.  The CPU oracle stands in for the
device engine, so this checks the *plumbing* of gr_raytracer_b200/renderer.py: the 2-rank image must be
bit-identical to the 1-rank image."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gr_raytracer_b200 import abi
from gr_raytracer_b200.renderer import FrameRenderer, tile_interleaved_items
from oracle import scene as osc
from helpers import ROOT, CONFIGS, config_opts

W, H = 40, 28


class OracleEngine:
    """DeviceEngine's interface on CPU tensors, backed by the oracle library (test infrastructure)."""

    def __init__(self, scene):
        self.sc = scene
        lib = osc.lib()
        u32p = C.POINTER(C.c_uint32)
        lib.orc_trace_items.argtypes = [C.POINTER(abi.GrrtScene), u32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_uint32, C.c_uint32, abi.c_double_p, abi.c_uint8_p, abi.c_uint8_p, u32p]
        self.lib = lib

    @staticmethod
    def _p(t, typ):
        return None if t is None else C.cast(t.data_ptr(), typ)

    def new(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def trace(self, items, n_items, rect, spa, xyza, cls=None, stop=None, steps=None):
        fr, fc, tr, tc = rect
        u32p = C.POINTER(C.c_uint32)
        self.lib.orc_trace_items(self.sc.ptr, self._p(items, u32p), 0, n_items, fr, fc, tr, tc, spa, self._p(xyza, abi.c_double_p),
                                 self._p(cls, abi.c_uint8_p), self._p(stop, abi.c_uint8_p), self._p(steps, u32p))

    def select(self, xyza, cls, rows, cols, selected, count):
        ml = C.c_double()
        self.lib.orc_select_mask(C.byref(self.sc.c.adaptive), self._p(xyza, abi.c_double_p), self._p(cls, abi.c_uint8_p), rows, cols,
                                 self._p(selected, C.POINTER(C.c_uint32)), self._p(count, C.POINTER(C.c_uint32)), C.byref(ml))
        return ml.value


def _scene():
    return osc.scene_from_toml(os.path.join(ROOT, CONFIGS["C2"]["toml"]), config_opts("C2", width=W, height=H), ROOT)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = _scene()
    fr = FrameRenderer(OracleEngine(sc), sc.c.adaptive, W, H, rank, world)
    info = fr.render()
    if rank == 0:
        np.savez(out_path, xyza=fr.xyza.numpy(), cls=fr.cls.numpy(), stop=fr.stop.numpy(), n_sel=info["pixels_supersampled"])
    dist.barrier()
    dist.destroy_process_group()


def test_tile_partition_is_a_partition():
    for (w, h, n) in ((40, 28, 2), (37, 23, 4), (1500, 1500, 8), (8, 4, 8)):
        parts = [tile_interleaved_items(w, h, r, n) for r in range(n)]
        allp = np.concatenate(parts)
        assert allp.size == w * h and np.array_equal(np.sort(allp), np.arange(w * h, dtype=np.uint32))
        sizes = [p.size for p in parts]
        assert max(sizes) - min(sizes) <= 32 + 0.02 * (w * h / n)   # partial edge tiles land on fixed ranks


def test_two_ranks_render_the_same_image_as_one(tmp_path):
    sc = _scene()
    single = FrameRenderer(OracleEngine(sc), sc.c.adaptive, W, H)
    info = single.render()
    assert info["pixels_supersampled"] > 0
    ref = osc.render(sc)  # and the renderer's 1-rank path equals the oracle's own frame driver
    assert np.array_equal(single.xyza.numpy().reshape(H, W, 4), ref["xyza"])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    assert int(got["n_sel"]) == info["pixels_supersampled"]
    assert np.array_equal(got["xyza"], single.xyza.numpy())
    assert np.array_equal(got["cls"], single.cls.numpy()) and np.array_equal(got["stop"], single.stop.numpy())


def test_section_render_and_cost_order_on_the_cpu_engine():
    """FrameRenderer on a section (the reference's --from-row ... --to-col) equals the oracle's own section render, and
    the cost order it derives is the stable descending-step-count order the device sort produces."""
    sc = _scene()
    rect = (4, 6, 24, 38)
    fr = FrameRenderer(OracleEngine(sc), sc.c.adaptive, W, H, rect=rect)
    info = fr.render()
    ref = osc.render(sc, *rect)
    assert np.array_equal(fr.xyza.numpy().reshape(20, 32, 4), ref["xyza"])
    assert info["pixels_supersampled"] == ref["stats"]["pixels_supersampled"] > 0
    n_sel = info["pixels_supersampled"]
    sel = fr.selected[:n_sel].numpy().astype(np.int64)
    order = fr.ordered[:n_sel].numpy().astype(np.int64)
    steps = fr.steps.numpy().astype(np.int64)
    assert np.array_equal(np.sort(order), sel)                               # a permutation of the flagged pixels
    want = sel[np.argsort(-steps[sel], kind="stable")]                       # ties keep ascending pixel order
    assert np.array_equal(order, want)
