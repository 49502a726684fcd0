"""bench.py's host-side logic without a GPU: the per-scene block's decisions (full frame / half frame / band of rows /
skipped / failing scene) with a stand-in for the device, and the bounded CPU sample of the reference arm."""
import importlib.util
import os
import time

from helpers import ROOT

spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench_py = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench_py)


class FakeBench:
    """Stands in for bench.Bench: every 'render' takes `cost[name]` seconds of pretended time."""

    def __init__(self, world, cost, fail=()):
        self.world, self.rank, self.cost, self.fail = world, 0, cost, set(fail)
        self.peak_tflops = 34.0
        self.ctx = object()
        self.calls = []
        self.clock = 0.0

        class _Api:
            @staticmethod
            def Context(devs):
                class _G:
                    def close(self_inner):
                        pass
                return _G()
        self.api = _Api()

    def reduce(self, values, op):
        return list(values)

    def host_barrier(self):
        pass

    def time_scene_cabi(self, group, wl, frames, warmup, rect=None):
        name = [k for k, v in bench_py.WORKLOADS.items() if v["toml"] == wl["toml"]][0]
        self.calls.append((name, rect, frames))
        if name in self.fail:
            raise RuntimeError("device lost")
        frac = 1.0
        if rect:
            H, W = wl["opts"]["height"], wl["opts"]["width"]
            frac = (rect[2] - rect[0]) * (rect[3] - rect[1]) / (H * W)
        self.clock += self.cost[name] * frac
        return dict(ms_per_frame=1e3 * self.cost[name] * frac, mrays=1.0, rays_per_frame=1000, pixels_supersampled=10,
                    steps_per_s=1e9, attempts_per_s=2e9, samples_per_s=0.0, stop_histogram=[0, 1, 2, 0, 0, 0], ray_errors=0,
                    tflops=10.0, base_ms=1.0, supersample_ms=1.0)


def run_block(monkeypatch, world, budget, cost, fail=()):
    fb = FakeBench(world, cost, fail)
    t0 = time.perf_counter()
    monkeypatch.setattr(bench_py.time, "perf_counter", lambda: t0 + fb.clock)   # pretended time
    return fb, bench_py.per_scene_block(fb, budget)


def test_per_scene_block_on_one_gpu(monkeypatch):
    cost = {"C1": 0.3, "C3": 160.0, "C4": 340.0, "C5": 83.0}
    fb, out = run_block(monkeypatch, 1, 420.0, cost)
    assert set(out) == {"C1", "C3", "C4", "C5"}
    assert out["C1"]["full_frame"] and out["C3"]["full_frame"]
    assert out["C4"]["full_frame"] is False and out["C4"]["section"]["to_col"] == 1024      # left half of the Kerr-Schild frame
    assert "section" in out["C5"] or out["C5"].get("full_frame") or "skipped" in out["C5"]
    assert fb.clock <= 420.0 * 1.1
    for name, e in out.items():
        if "ms_per_frame" in e:
            assert e["roofline_frac"] == 10.0 / 34.0 and e["parity"] and e["api"].startswith("grrt_render")


def test_per_scene_block_budgets_and_failures(monkeypatch):
    cost = {"C1": 0.3, "C3": 160.0, "C4": 340.0, "C5": 83.0}
    fb, out = run_block(monkeypatch, 1, 40.0, cost)
    assert out["C1"]["full_frame"] and "skipped" in out["C3"] and "skipped" in out["C4"]
    assert out["C5"]["full_frame"] is False and out["C5"]["section"]["from_col"] == 0       # a centred band of rows
    fb, out = run_block(monkeypatch, 8, 420.0, {"C1": 0.07, "C3": 111.0, "C4": 111.0, "C5": 10.7})
    assert all(out[k]["full_frame"] for k in out)                                           # eight GPUs: everything at full size
    assert "8 GPUs" in out["C3"]["api"]
    fb, out = run_block(monkeypatch, 1, 420.0, cost, fail=("C3",))
    assert "error" in out["C3"] and "device lost" in out["C3"]["error"] and out["C1"]["full_frame"]
    fb, out = run_block(monkeypatch, 1, 0.5, cost)
    assert all("skipped" in out[k] for k in ("C3", "C4", "C5"))


def test_strips_are_a_bounded_sample():
    s = bench_py.strips(1500, 25, 4)
    assert len(s) == 25 and all(b - a == 4 and 0 <= a and b <= 1500 for a, b in s)
    assert s[0][0] == 0 and s[-1][1] == 1500
    assert bench_py.strips(8, 100, 4) == [(0, 4), (4, 8)]
    assert bench_py.algorithmic_flops(4, 10, 5, 0) == 10 * 858 + 5 * 90
