"""Renders one BASELINE.json configuration at its stated size through the C ABI and checks a sample of its pixels
against the CPU oracle (test infrastructure; the render itself never touches it).

    python tests/fullsize_check.py C3 [--samples 300] [--long 30] [--out gpurun_out/fullsize_C3.json] [--width W --height H]

The frame is the configuration as given (adaptive supersampling included).  Checked per sampled pixel:
 * step count of the 1-spp ray bit-equal to the oracle's, same stop reason;
 * XYZ within 1e-4 relative of the oracle's 1-spp colour, or -- when the pixel was supersampled -- of the oracle's
   mean over the same 4x4 strata; ray class equal for 1-spp pixels.
`--long N` forces N of the samples to be rays with more than 1e5 accepted steps (shadow rays of the Kerr-Schild
scenes at --max-steps 1e6).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # helpers.py

import numpy as np  # noqa: E402


def oracle_pixels(osc, oscene, W, H, pids, spa):
    """orc_trace_items on a list of pixel ids: spa = 0 -> 1-spp results per item; spa > 0 -> mean over the strata."""
    n = len(pids)
    items = np.ascontiguousarray(pids, dtype=np.uint32)
    if spa == 0:
        xyza = np.zeros((n, 4))
        cls = np.zeros(n, dtype=np.uint8)
        stop = np.zeros(n, dtype=np.uint8)
        steps = np.zeros(n, dtype=np.uint32)
        osc.lib().orc_trace_items(C.byref(oscene.c), items.ctypes.data_as(C.POINTER(C.c_uint32)), 0, n, 0, 0, H, W, 0,
                                  osc.dptr(xyza), osc.u8ptr(cls), osc.u8ptr(stop), steps.ctypes.data_as(C.POINTER(C.c_uint32)))
        return xyza, cls, stop, steps
    frame = np.full((W * H, 4), np.nan)   # the supersampling form writes at 4 * pid
    osc.lib().orc_trace_items(C.byref(oscene.c), items.ctypes.data_as(C.POINTER(C.c_uint32)), 0, n, 0, 0, H, W, spa,
                              osc.dptr(frame), None, None, None)
    return frame[items], None, None, None


def rel_err(a, b):
    a = np.asarray(a)[..., :3]
    b = np.asarray(b)[..., :3]
    scale = np.max(np.abs(b), axis=-1)
    err = np.max(np.abs(a - b), axis=-1)
    return np.where(scale > 0, err / np.maximum(scale, 1e-300), err)


def run(name, n_samples=300, n_long=30, width=None, height=None, seed=2026, ctx=None, verbose=True, adaptive=True, devices=1):
    from gr_raytracer_b200 import abi, api
    from oracle import scene as osc
    from helpers import CONFIGS, config_opts
    kw = {}
    if width:
        kw["width"] = width
    if height:
        kw["height"] = height
    opts = config_opts(name, **kw)
    toml = os.path.join(ROOT, CONFIGS[name]["toml"])
    is_text = False
    if not adaptive:   # the 1-spp pass only
        toml, is_text = open(toml).read() + "\n[adaptive_sampling]\nenabled = false\n", True
    scene = api.HostScene(toml, opts, ROOT, is_text=is_text)
    own = ctx is None
    if own:
        ctx = api.Context(0) if devices <= 1 else api.Context(list(range(devices)))   # one multi-device context (grrt_init_multi)
    t0 = time.perf_counter()
    dev = ctx.render(scene)
    wall = time.perf_counter() - t0
    st = dev["stats"]
    H, W = dev["stop"].shape
    if verbose:
        print(f"[fullsize] {name} {W}x{H}: wall {wall:.2f} s, device {st['device_ms_total']:.0f} ms "
              f"(base {st['device_ms_base']:.0f}, supersample {st['device_ms_supersample']:.0f}); rays {st['rays_base']}+"
              f"{st['rays_supersample']}, steps {st['accepted_steps']:.4e}, attempts {st['attempts']:.4e}, "
              f"samples {st['raymarch_samples']:.4e}, stop histogram {st['stop_histogram']}", flush=True)
    # ---- sample ----
    rng = np.random.default_rng(seed)
    steps = dev["steps"].reshape(-1)
    pids = set()
    long_ids = np.flatnonzero(steps > 100000)
    if n_long and len(long_ids):
        pids.update(int(p) for p in rng.choice(long_ids, min(n_long, len(long_ids)), replace=False))
    n_long_got = len(pids)
    # a third from the central band (shadow / disc), the rest uniform
    while len(pids) < n_long_got + (n_samples - n_long_got) // 3:
        r = int(rng.integers(int(0.35 * H), int(0.65 * H)))
        c = int(rng.integers(int(0.35 * W), int(0.65 * W)))
        if steps[r * W + c] <= 100000 or name not in ("C3", "C4"):
            pids.add(r * W + c)
    while len(pids) < n_samples:
        p = int(rng.integers(0, W * H))
        if steps[p] <= 100000 or name not in ("C3", "C4"):   # long rays only through --long (CPU time)
            pids.add(p)
    pids = np.array(sorted(pids), dtype=np.int64)
    oscene = osc.scene_from_toml(toml, opts, ROOT, is_text=is_text)
    spa = int(oscene.c.adaptive.samples_per_axis) if oscene.c.adaptive.enabled else 0
    t0 = time.perf_counter()
    o_xyza, o_cls, o_stop, o_steps = oracle_pixels(osc, oscene, W, H, pids, 0)
    t_oracle = time.perf_counter() - t0
    d_xyza = dev["xyza"].reshape(-1, 4)[pids]
    d_cls = dev["ray_class"].reshape(-1)[pids]
    d_stop = dev["stop"].reshape(-1)[pids]
    d_steps = steps[pids]
    e1 = rel_err(d_xyza, o_xyza)
    fate = (d_stop == o_stop) & (d_steps == o_steps)
    ok1 = (e1 <= 1e-4) & fate & (d_cls == o_cls)
    # pixels that do not match the 1-spp colour may have been supersampled: compare with the oracle's strata mean
    ss_checked = 0
    ok = ok1.copy()
    err = e1.copy()
    if spa and not np.all(ok1):
        idx = np.flatnonzero(~ok1 & fate)
        ss_checked = len(idx)
        if ss_checked:
            m_xyza, _, _, _ = oracle_pixels(osc, oscene, W, H, pids[idx], spa)
            e2 = rel_err(d_xyza[idx], m_xyza)
            a_ok = np.abs(d_xyza[idx, 3] - m_xyza[:, 3]) <= 1e-4 * np.maximum(np.abs(m_xyza[:, 3]), 1e-300)
            ok[idx] = (e2 <= 1e-4) & a_ok
            err[idx] = e2
    bad = np.flatnonzero(~ok)
    res = {
        "config": name, "adaptive": bool(adaptive), "devices": ctx.device_count, "width": W, "height": H, "wall_s": wall, "device_ms_total": st["device_ms_total"],
        "device_ms_base": st["device_ms_base"], "device_ms_supersample": st["device_ms_supersample"],
        "rays_base": st["rays_base"], "rays_supersample": st["rays_supersample"], "pixels_supersampled": st["pixels_supersampled"],
        "accepted_steps": st["accepted_steps"], "attempts": st["attempts"], "raymarch_samples": st["raymarch_samples"],
        "stop_histogram": list(st["stop_histogram"]), "ray_errors": st["ray_errors"],
        "hit_stack_overflows": st.get("hit_stack_overflows"),
        "mrays_per_s": (st["rays_base"] + st["rays_supersample"]) / (st["device_ms_total"] * 1e-3) / 1e6,
        "steps_per_s": st["accepted_steps"] / (st["device_ms_total"] * 1e-3),
        "attempts_per_s": st["attempts"] / (st["device_ms_total"] * 1e-3),
        "samples_per_s": st["raymarch_samples"] / (st["device_ms_total"] * 1e-3),
        "max_steps_per_ray": int(steps.max()), "rays_over_1e5_steps": int((steps > 100000).sum()),
        "parity": {"sampled": int(len(pids)), "long_rays_sampled": int((d_steps > 100000).sum()),
                   "fate_equal": int(fate.sum()), "steps_bit_equal": int((d_steps == o_steps).sum()),
                   "within_gate": int(ok.sum()), "fraction": float(ok.mean()), "worst_rel_err": float(err[ok].max()) if ok.any() else None,
                   "checked_against_strata_mean": int(ss_checked), "oracle_seconds": t_oracle,
                   "outliers": [{"row": int(pids[i] // W), "col": int(pids[i] % W), "rel_err": float(err[i]),
                                 "dev_stop": int(d_stop[i]), "ora_stop": int(o_stop[i]), "dev_steps": int(d_steps[i]),
                                 "ora_steps": int(o_steps[i])} for i in bad[:50]]},
    }
    if verbose:
        p = res["parity"]
        print(f"[fullsize] {name}: {p['sampled']} pixels sampled ({p['long_rays_sampled']} with > 1e5 steps), fate equal "
              f"{p['fate_equal']}, steps bit-equal {p['steps_bit_equal']}, within 1e-4 {p['within_gate']} "
              f"({100 * p['fraction']:.2f} %), worst in-gate rel err {p['worst_rel_err']}, {p['checked_against_strata_mean']} "
              f"compared with the 4x4 strata mean, oracle {t_oracle:.1f} s, outliers {p['outliers']}", flush=True)
    scene.close()
    if own:
        ctx.close()
    return res, dev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--samples", type=int, default=300)
    ap.add_argument("--long", type=int, default=30)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--save-steps", default=None, help="write the per-pixel step counts and stop reasons (npz)")
    ap.add_argument("--base-only", action="store_true", help="1-spp pass only (adaptive supersampling off)")
    ap.add_argument("--devices", type=int, default=1, help="render through one multi-device context over this many GPUs")
    a = ap.parse_args()
    res, dev = run(a.config, a.samples, a.long, a.width, a.height, adaptive=not a.base_only, devices=a.devices)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
    if a.save_steps:
        np.savez_compressed(a.save_steps, steps=dev["steps"], stop=dev["stop"])
    print(json.dumps({k: v for k, v in res.items() if k != "parity"}))


if __name__ == "__main__":
    main()
