"""bench.py's host logic without a GPU: main() runs on stand-ins for the rasterizer and for torch's CUDA primitives, so that
an error in the argument handling, the timing loop or the assembly of the JSON line (the contract the driver parses)
shows up in the CPU suite instead of at round end.  Nothing is measured here; the numbers are meaningless."""
import io
import json
import os
import sys
import time
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = ("preprocess_fwd", "tile_scan", "emit", "tile_sort", "render_fwd", "render_bwd", "preprocess_bwd")


class _FakeEvent:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class _FakeRasterizer:
    def __init__(self, device):
        self.calls = 0

    def reserve(self, *a):
        pass

    def forward(self, cam, params, img=None, radii=None, defer_check=False):
        self.calls += 1
        return img, radii

    def backward(self, dl, grads, **kw):
        self.calls += 1

    def step_host(self, *a, **kw):
        self.calls += 1

    def stage_ms(self):
        return {k: 0.1 + 0.01 * i for i, k in enumerate(STAGES)}

    def stats(self):
        return {"num_visible": 8000, "num_dups": 60000, "tiles_x": 16, "tiles_y": 16, "max_tile_len": 400}

    def close(self):
        pass


@pytest.mark.skipif(torch.cuda.is_available(), reason="the real bench runs on a GPU box")
def test_bench_main_assembles_the_contract_line(monkeypatch, capfd):
    sys.path.insert(0, ROOT)
    import bench
    from divshot_b200 import rasterizer
    cpu = torch.device("cpu")
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch, "device", lambda *a, **k: cpu)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(rasterizer, "Rasterizer", _FakeRasterizer)
    monkeypatch.setattr(rasterizer, "scene_to_device", lambda sc, dev: {"means3D": torch.from_numpy(sc.means3D)})
    monkeypatch.setattr(bench, "measure_row", lambda tool: {"error": "not run in the mock: " + tool})
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "3", "--warmup", "3", "--no-cpu", "--workload", "c1"])
    monkeypatch.delenv("RANK", raising=False); monkeypatch.delenv("WORLD_SIZE", raising=False)
    bench.main()
    out = capfd.readouterr().out.strip().splitlines()
    rows = [l for l in out if l.startswith("{")]
    assert len(rows) == 1, out
    line = json.loads(rows[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "stages", "other_rows"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["gpu_launches"] == 30 and line["vs_baseline"] is None
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(line["other_rows"]) == {"F3_viewer_pack", "F1_refinement"}
    assert "workload" in line["config"] and line["config"]["workload"].startswith("c1")


def test_row_harnesses_report_failure_as_text_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sys.path.insert(0, ROOT)
    import bench
    for tool in ("bench_viewer_pack.py", "bench_densify.py"):
        r = bench.measure_row(tool)
        assert set(r) == {"error"} and "needs a GPU" in r["error"]
