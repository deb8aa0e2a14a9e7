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

    def step_host_async(self, *a, **kw):
        self.calls += 1

    def step_host_wait(self, slot):
        pass

    def set_profiling(self, on):
        pass

    def kernel_launches(self):  # 6 forward + 2 backward kernels per step, like the library's counter at c3
        return 4 * self.calls

    def stage_ms(self):
        return {k: 0.1 + 0.01 * i for i, k in enumerate(STAGES)}

    def stats(self):
        return {"num_visible": 8000, "num_dups": 60000, "tiles_x": 16, "tiles_y": 16, "max_tile_len": 400, "num_list_entries": 40000}

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
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["gpu_launches"] == 24 and line["vs_baseline"] is None
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(line["other_rows"]) == {"F3_viewer_pack", "F1_refinement", "F1_train_step"}
    assert "workload" in line["config"] and line["config"]["workload"].startswith("c1")


def test_row_harnesses_report_failure_as_text_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sys.path.insert(0, ROOT)
    import bench
    for tool in ("bench_viewer_pack.py", "bench_densify.py", "bench_trainstep.py"):
        r = bench.measure_row(tool)
        assert set(r) == {"error"} and "needs a GPU" in r["error"]


def _bench_rank(rank, world, port, q):
    """One rank of `bench.py --gpus 2` on stand-ins: gloo instead of NCCL, CPU tensors, the fake rasterizer."""
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from divshot_b200 import rasterizer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    cpu = torch.device("cpu")
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = _FakeEvent
    torch.device = lambda *a, **k: cpu
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **kw: real_init("gloo", rank=rank, world_size=world)

    class R(_FakeRasterizer):
        def backward(self, dl, grads, **kw):
            grads.flat.fill_(float(rank + 1))

        def step_host(self, cam, params, grads, *a, **kw):
            grads.flat.fill_(float(rank + 1))

        def step_host_async(self, cam, params, grads, *a, **kw):
            grads.flat.fill_(float(rank + 1))

    rasterizer.Rasterizer = R
    rasterizer.scene_to_device = lambda sc, dev: {"means3D": torch.from_numpy(sc.means3D)}
    bench.measure_row = lambda tool: {"error": "mock"}
    sys.argv = ["bench.py", "--gpus", str(world), "--steps", "2", "--warmup", "3", "--no-cpu", "--workload", "c2"]  # SH degree 1
    r_fd, w_fd = os.pipe()
    os.dup2(w_fd, 1)  # the JSON line goes to fd 1
    bench.main()
    sys.stdout.flush()
    os.set_blocking(r_fd, False)  # bench keeps its own dup of fd 1 open: do not wait for an end-of-file
    try:
        text = os.read(r_fd, 1 << 20).decode()
    except BlockingIOError:
        text = ""
    q.put((rank, text))


def test_bench_two_ranks_on_stand_ins_print_one_line_and_keep_the_plain_exchange():
    """N = 2 through main(): rank 0 alone prints the line; the factored exchange is set up, cannot run without CUDA, is voted
    down on both ranks, and the plain all-reduce carries the run (allreduce.backend / note say so)."""
    if torch.cuda.is_available():
        pytest.skip("the real bench runs on a GPU box")
    import socket

    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bench_rank, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    rows0 = [l for l in res[0].splitlines() if l.startswith("{")]
    rows1 = [l for l in res[1].splitlines() if l.startswith("{")]
    assert len(rows0) == 1 and not rows1
    line = json.loads(rows0[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["allreduce"]["backend"] == "nccl"
    assert "rejected" in line["allreduce"]["note"] and line["allreduce"]["ms"] > 0
    assert line["gpu_launches"] == 2 * 4 and "other_rows" not in line  # (the stand-in counts its forward only)
