"""Frame driver on device-resident buffers: the two-pass adaptive render of
`Raytracer::render_section_to_cie_buffer_supersampled` (reference src/rendering/raytracer.rs:257-318)
for one GPU or for the ranks of a `torch.distributed` job.

Multi-GPU (BASELINE.json north_star): pixels are independent, so the image is cut into 8x4-pixel
tiles dealt round-robin to the ranks (fine-grained interleaving balances the spatially clustered
shadow / disc work).  The only coupling is the adaptive mask: it reads the 8 neighbours of every
pixel and a frame-global 99th-percentile luminance (raytracer.rs:118-129, 386-458), so every rank
needs the whole 1-spp base buffer (37 B per pixel with the step counts); every rank then derives the
same mask, orders the flagged pixels by descending base-pass step count (the base pass is the pilot
of the second pass) and supersamples every world-th pixel of that list.  No reduction is involved;
the N-GPU image is bit-identical to the 1-GPU image.  The exchange is fused into the kernels: the
ranks map each other's frame buffers (cudaIpc, NVLink peer access) and the trace kernel / the
sub-sample mean kernel store every finished pixel into all of them, with a one-word all-reduce as
the barrier between passes and at the start of a frame.  Where the buffers cannot be shared (or with
GRRT_P2P=0) the same frame is assembled with NCCL all-gathers.  (One process driving all GPUs --
grrt_init_multi behind grrt_render -- is the C++ twin of this file, csrc/grrt_api.cu.)

torch is plumbing here (device buffers, the current stream, the collective); the ray tracing, mask
selection and tone mapping are the CUDA kernels of libgrrt.so reached through the C ABI.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import abi

TILE_W, TILE_H = 8, 4


def tile_interleaved_items(width, height, rank, world):
    """Pixel ids (row-major within the frame) of the 8x4 tiles owned by `rank`, tile-major so that 32
    consecutive work units are one compact patch."""
    tiles_x = (width + TILE_W - 1) // TILE_W
    tiles_y = (height + TILE_H - 1) // TILE_H
    tiles = np.arange(rank, tiles_x * tiles_y, world, dtype=np.int64)
    ty, tx = tiles // tiles_x, tiles % tiles_x
    dy, dx = np.divmod(np.arange(TILE_W * TILE_H, dtype=np.int64), TILE_W)
    py = (ty[:, None] * TILE_H + dy[None, :]).ravel()
    px = (tx[:, None] * TILE_W + dx[None, :]).ravel()
    ok = (py < height) & (px < width)
    return (py[ok] * width + px[ok]).astype(np.uint32)


class DeviceEngine:
    """libgrrt.so on one CUDA device, operating on torch tensors."""

    def __init__(self, ctx, scene):
        self.ctx = ctx
        self.scene = scene
        self.dscene = ctx.upload(scene)
        self.device = torch.device("cuda", ctx.device)
        self.counters = torch.zeros(16, dtype=torch.int64, device=self.device)

    def reupload(self, scene=None):
        """Uploads the scene again (host -> device), replacing the resident copy: one frame's input transfer."""
        if scene is not None:
            self.scene = scene
        old = self.dscene
        self.dscene = self.ctx.upload(self.scene)
        if old is not None:
            self.ctx.free_scene(old)

    def close(self):
        if self.dscene is not None:
            self.ctx.free_scene(self.dscene)
            self.dscene = None

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _stream(self):
        # torch's default stream has handle 0, which the C ABI reads as "use the context's own stream":
        # name the legacy default stream explicitly (cudaStreamLegacy == 0x1) so that work is ordered with torch's
        h = torch.cuda.current_stream(self.device).cuda_stream
        return C.c_void_p(h if h else 1)

    def trace(self, items, n_items, rect, spa, xyza, cls=None, stop=None, steps=None):
        self.ctx.trace_items_device(self.dscene, self._p(items), 0, n_items, rect, spa, self._p(xyza), self._p(cls),
                                    self._p(stop), self._p(steps), self._p(self.counters), self._stream())

    def select(self, xyza, cls, rows, cols, selected, count):
        return self.ctx.select_mask_device(self.dscene, self._p(xyza), self._p(cls), rows, cols, self._p(selected),
                                           self._p(count), self._stream())

    def new(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # -- frame buffers mapped into every rank of the box (cudaIpc), for the gather-free multi-GPU path ----------
    supports_peers = True

    def shared_alloc(self, nbytes):
        return self.ctx.ipc_alloc(nbytes)          # (device pointer, 64-byte handle)

    def shared_open(self, handle):
        return self.ctx.ipc_open(handle)

    def shared_close(self, ptr):
        self.ctx.ipc_close(ptr)

    def shared_free(self, ptr):
        self.ctx.ipc_free(ptr)

    def wrap(self, ptr, shape, dtype):
        """A torch view of raw device memory (not owned by torch)."""
        typestr = {torch.float64: "<f8", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)
        return torch.as_tensor(raw, device=self.device)

    def trace_peers(self, items, n_items, rect, spa, xyza_ptrs, cls_ptrs, stop_ptrs, steps_ptrs=None, item_offset=0,
                    item_stride=1):
        self.ctx.trace_items_device_peers(self.dscene, self._p(items), n_items, rect, spa, xyza_ptrs, cls_ptrs, stop_ptrs,
                                          self._p(self.counters), self._stream(), steps_ptrs, item_offset, item_stride)

    def order_by_cost(self, selected, n_selected, steps, ordered):
        self.ctx.order_by_cost_device(self._p(selected), n_selected, self._p(steps), self._p(ordered), self._stream())


class FrameRenderer:
    """One frame = base pass + mask + supersampling, sharded over `world` ranks."""

    def __init__(self, engine, adaptive, width, height, rank=0, world=1, group=None, rect=None):
        """`rect` = (from_row, from_col, to_row, to_col): render that section of the width x height frame (the
        reference's --from-row ... --to-col, cli.rs:63-74); pixel ids and the mask are then section-relative."""
        self.e, self.ad = engine, adaptive
        self.rect = (0, 0, int(height), int(width)) if rect is None else tuple(int(v) for v in rect)
        self.W, self.H = self.rect[3] - self.rect[1], self.rect[2] - self.rect[0]
        self.rank, self.world, self.group = rank, world, group
        n = self.W * self.H
        self.n = n
        self.p2p = False
        self._own, self._opened = [], []
        if 1 < world <= abi.MAX_PEERS and getattr(engine, "supports_peers", False) and os.environ.get("GRRT_P2P", "1") != "0":
            self._setup_peer_frames()
        if not self.p2p:
            self.xyza = engine.new((n, 4), torch.float64)
            self.cls = engine.new((n,), torch.uint8)
            self.stop = engine.new((n,), torch.uint8)
            self.steps = engine.new((n,), torch.int32)
        self.selected = engine.new((n,), torch.int32)
        self.ordered = engine.new((n,), torch.int32)   # flagged pixels by descending base-pass step count
        self.count = engine.new((1,), torch.int32)
        if world > 1:
            ids = tile_interleaved_items(self.W, self.H, rank, world)
            self.my_items = torch.from_numpy(ids.astype(np.int32)).to(self.xyza.device)
            self.n_mine = int(ids.size)
            self.max_items = max(int(tile_interleaved_items(self.W, self.H, r, world).size) for r in range(world))
            # where every slot of the gathered [world, max_items] buffers goes in the frame (padding slots dropped)
            per_rank = [tile_interleaved_items(self.W, self.H, r, world).astype(np.int64) for r in range(world)]
            src = np.concatenate([r * self.max_items + np.arange(ids_r.size, dtype=np.int64) for r, ids_r in enumerate(per_rank)])
            self.gather_src = torch.from_numpy(src).to(self.xyza.device)
            self.gather_dst = torch.from_numpy(np.concatenate(per_rank)).to(self.xyza.device)
            m = self.max_items
        if world > 1 and not self.p2p:
            self.loc_xyza = engine.new((m, 4), torch.float64)
            self.loc_cls = engine.new((m,), torch.uint8)
            self.loc_stop = engine.new((m,), torch.uint8)
            self.loc_steps = engine.new((m,), torch.int32)
            self.gat_xyza = engine.new((world, m, 4), torch.float64)
            self.gat_meta = engine.new((world, m, 2), torch.uint8)
            self.gat_steps = engine.new((world, m), torch.int32)
        self.last = {}
        self.profile = False   # True: CUDA-event timestamps of the phases of render() in self.phase_ms (adds syncs)
        self.phase_ms = {}

    def _setup_peer_frames(self):
        """Frame buffers of every rank mapped into every rank (cudaIpc over NVLink): the trace kernel stores each
        finished pixel into all of them, so no gather / scatter pass is needed.  Falls back (on every rank) to the
        all-gather path if any rank cannot export or map the buffers."""
        import torch.distributed as dist
        e, n, world = self.e, self.n, self.world
        ok = 1
        sizes = (n * 32, n, n, n * 4)
        try:
            own = [e.shared_alloc(sz) for sz in sizes]
        except Exception:
            own, ok = [], 0
        self._own = [p for p, _ in own]
        blob = b"".join(h for _, h in own) if ok else bytes(256)
        mine = torch.tensor(list(blob), dtype=torch.uint8, device=e.device)
        allh = torch.empty((world, 256), dtype=torch.uint8, device=e.device)
        dist.all_gather_into_tensor(allh.view(-1), mine, group=self.group)
        flag = torch.tensor([ok], dtype=torch.int32, device=e.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        peers = [[None] * world for _ in range(4)]
        if int(flag.item()):
            allh = allh.cpu().numpy()
            try:
                for r in range(world):
                    for b in range(4):
                        if r == self.rank:
                            peers[b][r] = self._own[b]
                        else:
                            ptr = e.shared_open(bytes(allh[r, 64 * b:64 * (b + 1)]))
                            self._opened.append(ptr)
                            peers[b][r] = ptr
            except Exception:
                ok = 0
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=e.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if not int(flag.item()):
            self.close()
            return
        self.peer_xyza, self.peer_cls, self.peer_stop, self.peer_steps = peers
        self.xyza = e.wrap(self._own[0], (n, 4), torch.float64)
        self.cls = e.wrap(self._own[1], (n,), torch.uint8)
        self.stop = e.wrap(self._own[2], (n,), torch.uint8)
        self.steps = e.wrap(self._own[3], (n,), torch.int32)
        self._flag = torch.zeros(1, dtype=torch.int32, device=e.device)
        self.p2p = True

    def close(self):
        """Unmaps / frees the shared frame buffers (the tensors xyza / cls / stop must not be used afterwards)."""
        if self.p2p or self._own or self._opened:
            torch.cuda.synchronize()
            if self.world > 1 and self.p2p:
                import torch.distributed as dist
                dist.barrier(group=self.group)   # nobody unmaps while a peer may still store into the buffers
            for ptr in self._opened:
                self.e.shared_close(ptr)
            for ptr in self._own:
                self.e.shared_free(ptr)
            self._opened, self._own, self.p2p = [], [], False

    def _barrier(self):
        """Orders the passes across ranks on the current stream: completes once every rank's earlier work has."""
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)

    def _all_gather(self, out, inp):
        import torch.distributed as dist
        dist.all_gather_into_tensor(out, inp, group=self.group)

    def _order(self, n_sel):
        """Flagged pixels by descending base-pass step count (the base pass is the pilot of the second pass; stable, so
        every rank derives the same list).  Traced in this order the work queue ends on short rays, and dealing the list
        round-robin gives every rank the same share of the expensive pixels.  The image does not depend on the order."""
        e = self.e
        if hasattr(e, "order_by_cost"):
            e.order_by_cost(self.selected, n_sel, self.steps, self.ordered)
        else:   # engines without the device sort (the CPU stand-in of the tests): same stable order with torch
            sel = self.selected[:n_sel].long()
            idx = torch.sort(-self.steps[sel].long(), stable=True).indices
            self.ordered[:n_sel] = self.selected[:n_sel][idx]

    def render(self):
        """Renders the frame into self.xyza / cls / stop / steps (device; steps = accepted steps of the 1-spp ray).
        Returns ray counts."""
        e, ad = self.e, self.ad
        marks = []

        def mark(name):
            if self.profile:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        if self.world == 1:
            e.trace(None, self.n, self.rect, 0, self.xyza, self.cls, self.stop, self.steps)
        elif self.p2p:
            # no rank may store pixels of this frame into a peer that is still reading (or copying out) the previous one
            self._barrier()
            e.trace_peers(self.my_items, self.n_mine, self.rect, 0, self.peer_xyza, self.peer_cls, self.peer_stop, self.peer_steps)
            mark("base")
            self._barrier()
        else:
            e.trace(self.my_items, self.n_mine, self.rect, 0, self.loc_xyza, self.loc_cls, self.loc_stop, self.loc_steps)
            mark("base")
            meta = torch.stack((self.loc_cls, self.loc_stop), dim=1)
            self._all_gather(self.gat_xyza.view(-1, 4), self.loc_xyza)
            self._all_gather(self.gat_meta.view(-1, 2), meta)
            self._all_gather(self.gat_steps.view(-1), self.loc_steps)
            self.xyza.index_copy_(0, self.gather_dst, self.gat_xyza.view(-1, 4).index_select(0, self.gather_src))
            meta_all = self.gat_meta.view(-1, 2).index_select(0, self.gather_src)
            self.cls.index_copy_(0, self.gather_dst, meta_all[:, 0].contiguous())
            self.stop.index_copy_(0, self.gather_dst, meta_all[:, 1].contiguous())
            self.steps.index_copy_(0, self.gather_dst, self.gat_steps.view(-1).index_select(0, self.gather_src))
        mark("gather_base")
        n_sel = 0
        min_lum = 0.0
        if ad.enabled or ad.show_sampling_mask:
            min_lum = e.select(self.xyza, self.cls, self.H, self.W, self.selected, self.count)
            n_sel = int(self.count.item())
            if n_sel and not ad.show_sampling_mask:
                self._order(n_sel)
            mark("mask")
            if self.p2p:
                self._barrier()   # no rank may store sub-sample means into a frame whose mask is still being computed
            if n_sel and ad.show_sampling_mask:
                col = torch.tensor(list(ad.mask_xyza), dtype=torch.float64, device=self.xyza.device)
                self.xyza[self.selected[:n_sel].long()] = col
            elif n_sel:
                spa = int(ad.samples_per_axis)
                if self.world == 1:
                    e.trace(self.ordered, n_sel, self.rect, spa, self.xyza)
                elif self.p2p:
                    k = (n_sel - self.rank + self.world - 1) // self.world if self.rank < n_sel else 0
                    if k:   # every world-th pixel of the cost-ordered list, starting at this rank
                        e.trace_peers(self.ordered, k, self.rect, spa, self.peer_xyza, self.peer_cls, self.peer_stop,
                                      None, self.rank, self.world)
                    mark("supersample")
                    self._barrier()
                else:
                    # flagged pixels dealt round-robin in cost order; every rank writes its means into a frame-sized scratch
                    mine = self.ordered[:n_sel][self.rank::self.world].contiguous()
                    k = int(mine.numel())
                    per = (n_sel + self.world - 1) // self.world
                    loc = self.e.new((per, 4), torch.float64)
                    scratch = self.xyza.clone()
                    if k:
                        e.trace(mine, k, self.rect, spa, scratch)
                        loc[:k] = scratch[mine.long()]
                    mark("supersample")
                    gat = self.e.new((self.world, per, 4), torch.float64)
                    self._all_gather(gat.view(-1, 4), loc)
                    for r in range(self.world):
                        ids = self.ordered[:n_sel][r::self.world].long()
                        self.xyza.index_copy_(0, ids, gat[r, :ids.numel()])
        mark("end")
        if self.profile:
            torch.cuda.synchronize()
            self.phase_ms = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
        ss = int(ad.samples_per_axis) ** 2 if (n_sel and not ad.show_sampling_mask) else 0
        self.last = dict(rays_base=self.n, pixels_supersampled=n_sel, rays_supersample=n_sel * ss,
                         minimum_luminance=min_lum)
        return self.last
