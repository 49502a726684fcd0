#!/usr/bin/env python
"""bench.py -- rays/s and RKF45 steps/s of the geodesic path on N B200s, next to the CPU oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1..C5] [--impl reference]

One "step" is one full frame of the workload: 1-spp base pass, adaptive mask, 4x4 supersampling of
the flagged pixels (reference src/rendering/raytracer.rs:257-318).  Default workload: BASELINE.json
configs[1] (kerr-bl.toml, 1500x1500, KerrBL a=0.499, adaptive 4x4).  `value` is Mrays/s with the
scene resident in HBM; `e2e` goes through the C-ABI call grrt_render() with host buffers (scene
upload and result read-back inside the timed region).  `--impl reference` times the CPU oracle
(oracle/, a C++ restatement of the reference: the Rust reference cannot be built in this image) on
all host cores on a bounded sample of the same frame.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

# BASELINE.json configs with the command lines of SURVEY.md 8(d)
WORKLOADS = {
    "C1": dict(desc="scenes/schwarzschild.toml 500x500 camera -16,0,3.5 theta=-3.142 max-steps 20000",
               toml="scenes/schwarzschild.toml", opts=dict(width=500, height=500, camera_position=[-16.0, 0.0, 3.5], theta=-3.142)),
    "C2": dict(desc="scenes/kerr-bl.toml 1500x1500 KerrBL a=0.499 camera -10,0,-0.5 theta=-3.14159 adaptive 4x4",
               toml="scenes/kerr-bl.toml", opts=dict(width=1500, height=1500, camera_position=[-10.0, 0.0, -0.5], theta=-3.14159)),
    "C3": dict(desc="scenes/kerr.toml 1500x1500 Kerr-Schild a=0.499 camera -10,0,-0.5 theta=1.52 psi=-1.57 max-steps 1000000",
               toml="scenes/kerr.toml", opts=dict(width=1500, height=1500, camera_position=[-10.0, 0.0, -0.5], theta=1.52, psi=-1.57, max_steps=1000000)),
    "C4": dict(desc="scenes/kerr-volumetric-streaky.toml 2048x2048 camera -17,0,1.5 theta=1.659 psi=-1.57 max-steps 1000000",
               toml="scenes/kerr-volumetric-streaky.toml",
               opts=dict(width=2048, height=2048, camera_position=[-17.0, 0.0, 1.5], theta=1.659, psi=-1.57, max_steps=1000000)),
    "C5": dict(desc="scenes/schwarzschild-volumetric-dense.toml 8192x8192 camera -16,0,3.5 theta=-3.142",
               toml="scenes/schwarzschild-volumetric-dense.toml",
               opts=dict(width=8192, height=8192, camera_position=[-16.0, 0.0, 3.5], theta=-3.142)),
}
# algorithmic FP64 operations: per RKF45 attempt, per accepted step, per ray-march sample
FLOPS_ATTEMPT = {0: 498, 1: 726, 2: 726, 3: 10302, 4: 858}   # by GRRT_GEOM_*; Kerr = reference finite-difference mode
FLOPS_ACCEPTED = {0: 30, 1: 60, 2: 60, 3: 230, 4: 90}
FLOPS_SAMPLE = 1000


def ensure_resources():
    res = os.path.join(ROOT, "resources")
    if not os.path.exists(os.path.join(res, "celestial.png")):
        sys.path.insert(0, os.path.join(ROOT, "scenes"))
        import make_textures
        make_textures.main(res)


def algorithmic_flops(geom_kind, attempts, accepted, samples):
    return attempts * FLOPS_ATTEMPT[geom_kind] + accepted * FLOPS_ACCEPTED[geom_kind] + samples * FLOPS_SAMPLE


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report that instead of guessing
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------
def strips(height, n_strips, rows_per_strip):
    """Evenly spread horizontal strips: a bounded, representative sample of the frame for the CPU arm."""
    n_strips = max(1, min(n_strips, height // rows_per_strip))
    starts = [int(round(i * (height - rows_per_strip) / max(1, n_strips - 1))) for i in range(n_strips)] if n_strips > 1 else [(height - rows_per_strip) // 2]
    return [(s, min(height, s + rows_per_strip)) for s in starts]


def cpu_sample(osc, oscene, width, height, budget_s, threads_hint=None):
    """Times the oracle on strips of the frame sized for ~budget_s seconds.  Returns dict."""
    lib = osc.lib()
    cores = lib.orc_num_threads()
    # calibrate on one thin strip in the middle band of the frame
    t0 = time.perf_counter()
    r = osc.render(oscene, height // 3, 0, height // 3 + 2, width)
    dt = time.perf_counter() - t0
    rows = max(4, int(budget_s / max(dt / 2.0, 1e-4)))
    rows_per_strip = 4
    n_strips = max(1, rows // rows_per_strip)
    sel = strips(height, n_strips, rows_per_strip)
    rays = steps = att = samples = 0
    t0 = time.perf_counter()
    for (a, b) in sel:
        r = osc.render(oscene, a, 0, b, width)
        st = r["stats"]
        rays += st["rays_base"] + st["rays_supersample"]
        steps += st["accepted_steps"]
        att += st["attempts"]
        samples += st["raymarch_samples"]
    dt = time.perf_counter() - t0
    return dict(seconds=dt, rays=rays, steps=steps, attempts=att, samples=samples, cores=cores,
                sample=f"{len(sel)} strips x {rows_per_strip} rows x {width} cols of the {width}x{height} frame "
                       f"(base + mask + supersampling inside each strip)")


def run_reference(args, wl):
    """--impl reference: the CPU oracle on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gr_raytracer_b200 import abi
    from oracle import scene as osc
    ensure_resources()
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host core regardless
    osc.lib().orc_set_num_threads(len(os.sched_getaffinity(0)) or 1)
    opts = abi.default_cli_opts(**wl["opts"])
    oscene = osc.scene_from_toml(os.path.join(ROOT, wl["toml"]), opts, ROOT)
    W, H = wl["opts"]["width"], wl["opts"]["height"]
    total_budget = 150.0
    per_step = max(2.0, total_budget / (args.steps + args.warmup))
    res = None
    t_all = 0.0
    tot = dict(rays=0, steps=0, attempts=0, samples=0)
    for i in range(args.warmup + args.steps):
        res = cpu_sample(osc, oscene, W, H, per_step)
        if i >= args.warmup:
            t_all += res["seconds"]
            for k in tot:
                tot[k] += res[k]
    mrays = tot["rays"] / t_all / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "rkf45_steps_per_s": tot["steps"] / t_all, "rkf45_attempts_per_s": tot["attempts"] / t_all,
        "config": {"workload": args.workload + ": " + wl["desc"], "width": W, "height": H, "sampled": True,
                   "sample": res["sample"], "cores": res["cores"],
                   "note": "same scene, camera and command line as the GPU arm; each step traces a bounded sample of the frame "
                           "(evenly spread 4-row strips, base pass + mask + supersampling inside each strip) and the rate is "
                           "per ray, so it compares with the GPU arm's whole-frame rate"},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": res["cores"], "kind": "port", "sample": res["sample"],
                         "note": "C++ oracle (restatement of the Rust reference, OpenMP over pixels); the reference needs cargo, absent here"},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def scene_h2d_bytes(sc):
    """Bytes grrt_render copies to the device per call: textures (deduplicated) + LUTs + parameter block."""
    from gr_raytracer_b200 import abi
    seen, total = set(), 0
    texs = [sc.celestial] + [sc.objects[k].texture for k in range(sc.n_objects)]
    for t in texs:
        if t.kind == abi.TEX_BITMAP:
            key = C.addressof(t.rgba.contents)
            if key not in seen:
                seen.add(key)
                total += t.width * t.height * 4
    for k in range(sc.n_objects):
        o = sc.objects[k]
        if o.kind != abi.OBJ_SPHERE and o.temperature.kind == abi.TEMP_LUT:
            total += o.temperature.n * 16
        if o.kind == abi.OBJ_VOLUMETRIC_DISC:
            total += 256
    total += sc.bb_lut_n * 32
    return total + C.sizeof(abi.GrrtScene)


# expected seconds of one full frame (base + adaptive pass) on ONE B200, measured in round 2 (profiles/r2_fullsize_*.json):
# sizes the per-scene block so that the default run stays within minutes
EXPECTED_S_1GPU = {"C1": 0.3, "C2": 0.22, "C3": 165.0, "C4": 350.0, "C5": 85.0}
# ... and the floor more GPUs cannot lower: a Kerr-Schild ray that uses the whole 1e6-step budget is a ~47 s serial chain,
# and both passes of a frame contain such rays (8 GPUs: 111 s per frame for C3 and for C4)
FLOOR_S = {"C3": 105.0, "C4": 105.0}
PER_SCENE_ORDER = ["C1", "C3", "C4", "C5"]
FULL_FRAME_RECORD = {"C1": "tests/test_gpu_fullsize.py", "C3": "profiles/r2_fullsize_C3.json", "C4": "profiles/r2_fullsize_C4.json",
                     "C5": "profiles/r2_fullsize_C5_8192_1gpu.json"}


class Bench:
    """One process per GPU (torchrun for N > 1); frames are rendered device-resident by FrameRenderer."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from gr_raytracer_b200 import api
        self.torch, self.dist, self.api = torch, dist, api
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if args.gpus > 1 and self.world == 1:
            raise SystemExit("for --gpus N > 1 launch with torch.distributed.run (one rank per GPU)")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libgrrt has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.cpu_group = dist.new_group(backend="gloo")   # host-side waits that leave the GPUs idle
        if self.rank == 0:
            ensure_resources()
        if self.world > 1:
            dist.barrier()
        self.ctx = api.Context(self.local_rank)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        self.dfma_rate, _ = self.ctx.measure_fp64_peak()   # FP64 roofline denominator, measured here
        self.peak_tflops = 2.0 * self.dfma_rate / 1e12

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        """Waits on the CPU (gloo): unlike an NCCL barrier no kernel spins on the waiting ranks' GPUs."""
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(v) for v in t.tolist()]

    def time_frames(self, wl, steps, warmup, rect=None, want_clocks=False):
        """Renders warmup + steps frames of a workload (all ranks), device-timed per frame, max over ranks."""
        torch = self.torch
        from gr_raytracer_b200 import abi
        from gr_raytracer_b200.renderer import DeviceEngine, FrameRenderer
        opts = abi.default_cli_opts(**wl["opts"])
        scene = self.api.HostScene(os.path.join(ROOT, wl["toml"]), opts, ROOT)
        sc = scene.c
        W, H = int(sc.camera.cols), int(sc.camera.rows)
        engine = DeviceEngine(self.ctx, scene)
        fr = FrameRenderer(engine, sc.adaptive, W, H, self.rank, self.world, rect=rect)
        for _ in range(warmup):
            fr.render()
            self.flush.fill_(1)
        self.sync_all()
        engine.counters.zero_()
        launches0 = self.ctx.launch_count()
        sampler = None
        if want_clocks:
            sampler = ClockSampler(self.local_rank)
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        counts = None
        self.sync_all()
        for i in range(steps):
            ev[i][0].record()
            counts = fr.render()
            ev[i][1].record()
            self.flush.fill_(i & 1)       # L2 flush between timed iterations (not timed)
        self.sync_all()
        clocks = sampler.finish() if sampler else None
        launches = self.ctx.launch_count() - launches0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        ms = self.reduce([ms], "MAX")[0]
        cnt = engine.counters.clone()
        if self.world > 1:
            self.dist.all_reduce(cnt, op=self.dist.ReduceOp.SUM)
        launches = int(self.reduce([launches], "SUM")[0])
        cnt = [int(v) for v in cnt.tolist()]
        secs = ms * 1e-3
        rays = (counts["rays_base"] + counts["rays_supersample"]) * steps
        flops = algorithmic_flops(sc.geometry.kind, cnt[0], cnt[1], cnt[2])
        res = dict(scene=scene, engine=engine, fr=fr, W=W, H=H, ms_per_frame=ms / steps, secs=secs, rays_per_frame=rays // steps,
                   pixels_supersampled=counts["pixels_supersampled"], mrays=rays / secs / 1e6, steps_per_s=cnt[1] / secs,
                   attempts_per_s=cnt[0] / secs, samples_per_s=cnt[2] / secs, ray_errors=cnt[3],
                   stop_histogram=[c // steps for c in cnt[8:14]], flops=flops, tflops=flops / secs / 1e12,
                   launches=launches, clocks=clocks, p2p=fr.p2p)
        return res

    def time_scene_cabi(self, group, wl, frames, warmup, rect=None):
        """Frames of a workload through the reference-facing call (grrt_render with host buffers) on `group` -- this
        rank's context, or one grrt_init_multi context over all GPUs driven by rank 0 while the other ranks wait on the
        host.  Frame time = the library's CUDA-event time from the first kernel of the frame (pilot pass included) to
        the last, i.e. without the read-back of the result."""
        from gr_raytracer_b200 import abi
        self.sync_all()
        self.host_barrier()
        res = None
        if self.rank == 0:
            opts = abi.default_cli_opts(**wl["opts"])
            scene = self.api.HostScene(os.path.join(ROOT, wl["toml"]), opts, ROOT)
            kind = scene.c.geometry.kind
            ms = att = steps = samples = rays = 0
            out = None
            for i in range(warmup + frames):
                out = group.render(scene, *(rect or (None, None, None, None)))
                st = out["stats"]
                if i >= warmup:
                    ms += st["device_ms_total"]
                    att += st["attempts"]
                    steps += st["accepted_steps"]
                    samples += st["raymarch_samples"]
                    rays += st["rays_base"] + st["rays_supersample"]
            st = out["stats"]
            secs = ms * 1e-3
            flops = algorithmic_flops(kind, att, steps, samples)
            res = dict(ms_per_frame=ms / frames, mrays=rays / secs / 1e6, rays_per_frame=rays // frames,
                       pixels_supersampled=st["pixels_supersampled"], steps_per_s=steps / secs, attempts_per_s=att / secs,
                       samples_per_s=samples / secs, stop_histogram=list(st["stop_histogram"]), ray_errors=st["ray_errors"],
                       tflops=flops / secs / 1e12, base_ms=st["device_ms_base"], supersample_ms=st["device_ms_supersample"])
            scene.close()
        self.host_barrier()
        return res

    def close_frames(self, res):
        res["fr"].close()
        res["engine"].close()
        res["scene"].close()


def per_scene_block(bench, budget_s):
    """One frame of every other BASELINE.json configuration (the headline workload is timed separately), full size
    while the time budget lasts; a configuration that does not fit is rendered on a centred band of rows sized from
    the round-2 full-frame measurements (its rates are per ray / per step, hence comparable)."""
    out = {}
    world = bench.world
    group = None
    if bench.rank == 0:   # the drop-in call: one context over all GPUs of the job
        group = bench.ctx if world == 1 else bench.api.Context(list(range(world)))
    t_start = time.perf_counter()
    for name in PER_SCENE_ORDER:
        wl = WORKLOADS[name]
        W, H = wl["opts"]["width"], wl["opts"]["height"]
        left = budget_s - (time.perf_counter() - t_start)
        left = bench.reduce([left], "MIN")[0]                 # every rank takes the same decision
        est = max(EXPECTED_S_1GPU[name] / world, FLOOR_S.get(name, 0.0))
        expect = est * 1.2 + 3.0                              # + scene set-up, texture upload
        rect, full = None, True
        if expect > left and name in ("C3", "C4"):
            # Kerr-Schild frames are dominated by the shadow's crawling rays (up to 1e6 steps, ~20 s each), which sit in the
            # middle of the frame: the only representative section is a half frame cut through the shadow
            if est * 0.5 * 1.2 + 3.0 <= left:
                rect, full = (0, 0, H, W // 2), False
            else:
                out[name] = {"workload": wl["desc"], "skipped": f"time budget of the per-scene block used up ({budget_s:.0f} s); "
                             f"a full frame takes about {est:.0f} s on {world} GPU(s): {FULL_FRAME_RECORD[name]}"}
                continue
        elif expect > left:
            frac = (left - 3.0) / (est * 1.5)   # the centre rows are the expensive ones
            rows = int(H * frac) // 4 * 4
            if rows < 16:
                out[name] = {"workload": wl["desc"], "skipped": f"time budget of the per-scene block used up ({budget_s:.0f} s); "
                             f"a full frame takes about {est:.0f} s on {world} GPU(s): {FULL_FRAME_RECORD[name]}"}
                continue
            r0 = (H - rows) // 2
            rect, full = (r0, 0, r0 + rows, W), False
        frames = 3 if EXPECTED_S_1GPU[name] < 2.0 else 1
        try:
            r = bench.time_scene_cabi(group, wl, frames, 1 if frames > 1 else 0, rect=rect)
        except Exception as e:   # the headline line must survive a failing extra scene
            out[name] = {"workload": wl["desc"], "error": f"{type(e).__name__}: {e}"}
            bench.host_barrier()
            continue
        if r is None:            # ranks other than 0 only wait
            continue
        entry = {"workload": wl["desc"], "full_frame": full, "frames_timed": frames, "ms_per_frame": r["ms_per_frame"],
                 "mrays_per_s": r["mrays"], "rays_per_frame": r["rays_per_frame"], "pixels_supersampled": r["pixels_supersampled"],
                 "rkf45_steps_per_s": r["steps_per_s"], "rkf45_attempts_per_s": r["attempts_per_s"],
                 "raymarch_samples_per_s": r["samples_per_s"], "stop_histogram": r["stop_histogram"], "ray_errors": r["ray_errors"],
                 "algorithmic_tflops": r["tflops"], "roofline_frac": r["tflops"] / (bench.peak_tflops * world),
                 "base_pass_ms": r["base_ms"], "supersampling_pass_ms": r["supersample_ms"],
                 "api": "grrt_render (C ABI, host buffers)" if world == 1 else
                        f"grrt_render on one grrt_init_multi context over {world} GPUs (one process; the other ranks idle)",
                 "timing": "CUDA events inside the library, first kernel of the frame (pilot pass included) to the last",
                 "build": "exact (no FMA: the issue-slot ceiling is half the DFMA peak)" if name in ("C3", "C4") else "FMA",
                 "parity": PARITY_RECORD.get(name)}
        if rect:
            entry["section"] = {"from_row": rect[0], "to_row": rect[2], "from_col": rect[1], "to_col": rect[3],
                                "note": "section of the full-size frame (the reference CLI's --from-row/--to-row/--from-col/--to-col): the left "
                                        "half for the Kerr-Schild scenes, a centred band of rows otherwise; the whole frame is recorded in "
                                        "%s" % FULL_FRAME_RECORD[name]}
        out[name] = entry
    if group is not None and world > 1:
        group.close()
    return out


# where each configuration's parity at its stated size is established (tests/ run by `pytest -m gpu`; the oracle is
# not used by this file outside the cpu_baseline leg)
PARITY_RECORD = {
    "C1": "tests/test_gpu_fullsize.py::test_full_frame_parity_gate[C1]: 250000 of 250000 pixels of the 500x500 frame inside the gate",
    "C2": "tests/test_gpu_fullsize.py::test_full_frame_parity_gate[C2]: 99.999 % of the 1500x1500 frame inside the gate (21 outliers listed)",
    "C3": "tests/test_gpu_fullsize.py::test_kerr_schild_configs_at_stated_size[C3] + profiles/r2_fullsize_C3.json: 300 of 300 sampled pixels "
          "(30 shadow rays > 1e5 steps) bit-equal step counts, XYZ within 1e-12",
    "C4": "tests/test_gpu_fullsize.py::test_kerr_schild_configs_at_stated_size[C4] + profiles/r2_fullsize_C4.json: 300 of 300 sampled pixels",
    "C5": "profiles/r2_fullsize_C5_8192_1gpu.json / _8gpu_*.json: 300 of 300 sampled pixels of the 8192x8192 frame inside the gate with bit-equal step "
          "counts (1 GPU and one 8-device context) + tests/test_reference_image.py (Perlin pinned against the reference's own render)",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="grrt", choices=["grrt", "reference"])
    ap.add_argument("--width", type=int, default=None, help="override the workload's frame size (not a bench value)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-scene-budget-s", type=float, default=420.0,
                    help="seconds the per-scene block (one frame each of C1, C3, C4, C5) may take; 0 disables it")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    wl["opts"] = dict(wl["opts"])
    if args.width:
        wl["opts"]["width"] = args.width
    if args.height:
        wl["opts"]["height"] = args.height
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    from gr_raytracer_b200 import abi
    bench = Bench(args)
    rank, world, ctx, dist = bench.rank, bench.world, bench.ctx, bench.dist

    # ---- headline workload: K timed frames ----
    r = bench.time_frames(wl, args.steps, args.warmup, want_clocks=True)
    scene, engine, fr = r["scene"], r["engine"], r["fr"]
    sc = scene.c
    W, H = r["W"], r["H"]
    launches = r["launches"]

    # per-kernel time of the dominant kernel (trace_kernel): separate timed pass with events around each launch
    kern_ms, kern_flops = None, None
    if world == 1:
        engine.counters.zero_()
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        torch.cuda.synchronize()
        e0.record()
        engine.trace(None, fr.n, fr.rect, 0, fr.xyza, fr.cls, fr.stop, fr.steps)
        e1.record()
        torch.cuda.synchronize()
        n_sel = 0
        if sc.adaptive.enabled and not sc.adaptive.show_sampling_mask:
            engine.select(fr.xyza, fr.cls, H, W, fr.selected, fr.count)
            n_sel = int(fr.count.item())
        ss_ms = 0.0
        if n_sel:
            engine.order_by_cost(fr.selected, n_sel, fr.steps, fr.ordered)
            e2.record()
            engine.trace(fr.ordered, n_sel, fr.rect, int(sc.adaptive.samples_per_axis), fr.xyza)
            e3.record()
            torch.cuda.synchronize()
            ss_ms = e2.elapsed_time(e3)   # includes the (tiny) resolve kernel
        tot_cnt = [int(v) for v in engine.counters.tolist()]
        n_launch = 1 + (1 if n_sel else 0)
        kern_ms = (e0.elapsed_time(e1) + ss_ms) / n_launch
        kern_flops = algorithmic_flops(sc.geometry.kind, tot_cnt[0], tot_cnt[1], tot_cnt[2]) / n_launch

    # ---- N > 1: the N-rank frame must equal the frame one GPU renders alone ----
    bit_identical = None
    if world > 1:
        fr.render()
        bench.sync_all()
        if rank == 0:
            alone = ctx.render(scene)
            got = fr.xyza.cpu().numpy().reshape(H, W, 4)
            bit_identical = bool(np.array_equal(got, alone["xyza"]) and
                                 np.array_equal(fr.cls.cpu().numpy().reshape(H, W), alone["ray_class"]) and
                                 np.array_equal(fr.stop.cpu().numpy().reshape(H, W), alone["stop"]) and
                                 np.array_equal(fr.steps.cpu().numpy().reshape(H, W).view(np.uint32), alone["steps"]))
        bench.sync_all()

    # ---- end to end through the C ABI with HOST buffers: grrt_render uploads the scene, traces, reads the frame back ----
    # N > 1: rank 0 drives all N GPUs through ONE multi-device context (grrt_init_multi), the drop-in call a host program
    # makes; the other ranks wait at the barrier with their GPUs idle.
    e2e = None
    npx = W * H
    bench.sync_all()
    bench.host_barrier()
    if rank == 0:
        pinned = {"xyza": torch.empty((npx, 4), dtype=torch.float64).pin_memory().numpy(),
                  "ray_class": torch.empty(npx, dtype=torch.uint8).pin_memory().numpy(),
                  "stop": torch.empty(npx, dtype=torch.uint8).pin_memory().numpy(),
                  "steps": torch.empty(npx, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
        group = ctx if world == 1 else bench.api.Context(list(range(world)))
        for _ in range(2):
            out = group.render(scene, out=pinned)  # warm
        reps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(reps):
            out = group.render(scene, out=pinned)
        dt = (time.perf_counter() - t0) / reps
        st = out["stats"]
        e2e_rays = st["rays_base"] + st["rays_supersample"]
        d2h = W * H * (32 + 1 + 1 + 4) + 16 * 8 * world
        # one host -> device transfer of the scene per frame; the other devices of a multi-device context get the tables by
        # peer copies from the first
        e2e = {"value": e2e_rays / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": scene_h2d_bytes(sc),
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3,
               "api": "grrt_render (C ABI, host buffers)" if world == 1 else
                      f"grrt_render on one grrt_init_multi context over {world} GPUs (C ABI, host buffers, one process)"}
        if world > 1:
            e2e["bit_identical_to_1gpu"] = bool(np.array_equal(out["xyza"], alone["xyza"]) and np.array_equal(out["steps"], alone["steps"]))
            group.close()
    bench.host_barrier()
    headline = {k: r[k] for k in ("ms_per_frame", "mrays", "steps_per_s", "attempts_per_s", "samples_per_s", "ray_errors", "tflops",
                                  "rays_per_frame", "pixels_supersampled", "clocks", "p2p", "stop_histogram", "secs", "flops")}
    bench.close_frames(r)

    # ---- the other configurations ----
    per_scene = per_scene_block(bench, args.per_scene_budget_s) if (args.per_scene_budget_s > 0 and args.workload == "C2"
                                                                    and not args.width and not args.height) else None

    line = None
    if rank == 0:
        peak_tflops = bench.peak_tflops
        roofline = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak_tflops,
                    "peak_source": "DFMA-chain microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry); FMA = 2 flops",
                    "traffic": None}
        if args.workload == "C2" and (W, H) == (1500, 1500) and world == 1:
            roofline.update(C2_NCU_RECORD)
        if kern_ms:
            ach = kern_flops / (kern_ms * 1e-3) / 1e12
            roofline.update({"achieved": ach, "frac": ach / peak_tflops, "kernel": "trace_kernel", "kernel_ms_per_launch": kern_ms,
                             "algorithmic_flops_per_launch": kern_flops,
                             "note": "algorithmic FP64 ops of SURVEY.md 8(d) (divisions, sqrt and libm calls count 1 each; the reference's "
                                     "work on components nobody reads counts too) over the mean trace_kernel launch time; Kerr-Schild scenes "
                                     "run the contraction-free build (bit parity), whose issue-slot ceiling is half the FMA peak; the other "
                                     "geometries run the FMA build"})
        else:
            ach = headline["tflops"]
            roofline.update({"achieved": ach, "frac": ach / (peak_tflops * world), "kernel": "whole step, all ranks"})
        line = {
            "metric": "Mrays/s", "value": headline["mrays"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": headline["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "width": W, "height": H, "l2": "256 MiB flush write between timed steps",
                       "partition": "whole frame" if world == 1 else f"8x4 tiles round-robin over {world} ranks; flagged pixels dealt "
                                    "round-robin in descending base-pass step count",
                       "exchange": "none" if world == 1 else ("peer stores from the trace kernel (cudaIpc over NVLink)"
                                                              if headline["p2p"] else "NCCL all-gather"),
                       "textures": "synthetic 2048x2048 atlases (scenes/make_textures.py)"},
            "rays_per_step": headline["rays_per_frame"], "pixels_supersampled": headline["pixels_supersampled"],
            "rkf45_steps_per_s": headline["steps_per_s"], "rkf45_attempts_per_s": headline["attempts_per_s"],
            "raymarch_samples_per_s": headline["samples_per_s"], "stop_histogram": headline["stop_histogram"],
            "ray_errors": headline["ray_errors"], "algorithmic_tflops": headline["tflops"],
            "gpu_launches": launches, "clocks": headline["clocks"], "roofline": roofline, "e2e": e2e,
        }
        if bit_identical is not None:
            line["bit_identical_to_1gpu"] = bit_identical
        if per_scene is not None:
            line["per_scene"] = per_scene
    # CPU baseline on rank 0 at N=1 only (bounded sample, all host cores)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import scene as osc
        opts = abi.default_cli_opts(**wl["opts"])
        oscene = osc.scene_from_toml(os.path.join(ROOT, wl["toml"]), opts, ROOT)
        res = cpu_sample(osc, oscene, W, H, 20.0)
        line["cpu_baseline"] = {"value": res["rays"] / res["seconds"] / 1e6, "unit": "Mrays/s", "cores": res["cores"], "kind": "port",
                                "sample": res["sample"], "rkf45_steps_per_s": res["steps"] / res["seconds"], "seconds": res["seconds"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of the base-pass launch and the FP64-pipe figure from the ncu --set full capture
# of this command (C2 1500x1500, one GPU); see profiles/
C2_NCU_RECORD = {
    "traffic": 27.091456e6 + 81.335808e6,
    "traffic_source": "profiles/r2_c2_trace_kernel_inline_fallbacks_1500.json (ncu --set full, base-pass launch of this command: the 85.5 MB "
                      "of frame buffers written once, scene tables read once)",
    "fp64_pipe_busy_pct": 62.25,
    "fp64_pipe_source": "profiles/r2_c2_trace_kernel_final_1500.json (sm__inst_executed_pipe_fp64, base-pass launch, 1500x1500)",
}


import ctypes as C  # noqa: E402

if __name__ == "__main__":
    main()
