#!/usr/bin/env python
"""Measurement for row F3 (trainer -> viewer hand-off), same conventions as bench.py: one JSON line.

  python tools/bench_viewer_pack.py [--n 1000000] [--steps 50] [--warmup 5]        GPU kernel + end-to-end hand-off
  python tools/bench_viewer_pack.py --impl reference                               the reference's CPU quantiser only

value    = Gaussians/s of dvs_viewer_pack with parameters resident in HBM (CUDA events, L2 flushed between steps by the
           workload itself: 340 MB of traffic per step at 10^6 Gaussians > 126 MB L2)
roofline = 340 B/Gaussian algorithmic (236 read + 104 written) / event time vs the measured HBM peak
e2e      = the same through GaussianTrainerScene-style hand-off: pack kernel + D2H of the 104 B/Gaussian into pinned memory
cpu_baseline (kind "reference") = oracle/_ref/libviewerpack_ref.so, i.e. GaussianModel::create_gpu_buffer's own lines
           (gaussian_model.cpp:130-211) on the host cores (the reference runs them under its parallel_for: OpenMP here).
           Its real path also pays the 236 B/Gaussian D2H first (editor.cpp:1559-1566), not included.
STAGED: written in round 1 without a GPU; the CPU leg runs anywhere oracle/_ref exists."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import viewer_pack_util as u  # noqa: E402

BYTES_PER_GAUSSIAN = 236 + 104


def cpu_reference(n, reps=3):
    if not os.path.exists(u.REF_SO):
        return None
    m = u.make_model(n, 1)
    R = C.CDLL(u.REF_SO)
    aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best, cores = 1e30, 1
    for th in sorted({max(1, aff // d) for d in (1, 2, 4, 8)}, reverse=True):  # the reference runs this loop on its thread pool
        got = R.ref_viewer_pack_threads(th)
        for _ in range(reps):
            t = time.perf_counter(); u.pack_with_reference(m); dt = time.perf_counter() - t
            if dt < best:
                best, cores = dt, got
    return {"value": n / best, "unit": "Gaussians/s", "cores": cores, "kind": "reference",
            "sample": f"oracle/_ref/libviewerpack_ref.so (gaussian_model.cpp:130-211 compiled unmodified, its parallel_for as OpenMP), "
                      f"{n} Gaussians, best thread count of the affinity count and its 1/2, 1/4, 1/8, best of {reps} ({best:.3f} s)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    base = {"metric": "viewer hand-off Gaussians/s", "unit": "Gaussians/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "dtype": "u32/f16 records from f32", "data": "synthetic", "vs_baseline": None,
            "config": {"workload": f"F3: {a.n} Gaussians, SH degree 3 (236 B in, 104 B out per Gaussian)"}}
    if a.impl == "reference":
        cb = cpu_reference(min(a.n, 1000000))
        if cb is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libviewerpack_ref.so not built"}))
            return
        print(json.dumps({**base, "impl": "reference", "value": cb["value"], "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": "Gaussians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    from bench import _peaks
    from divshot_b200 import build
    assert torch.cuda.is_available(), "needs a GPU (no CPU path in the product)"
    torch.zeros(1, device="cuda")
    lib = C.CDLL(build.build_gstrain())
    lib.dvs_viewer_pack.argtypes = [C.c_void_p] * 6 + [C.c_int64] + [C.c_void_p] * 5
    n = a.n
    m = u.make_model(n, 1)
    t = {k: torch.from_numpy(np.ascontiguousarray(m[k])).cuda() for k in u.KEYS}
    g = torch.empty((n, 8), dtype=torch.int32, device="cuda"); c = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    sh = torch.empty((n, 16), dtype=torch.int32, device="cuda"); bb = torch.zeros(8, dtype=torch.int32, device="cuda")
    hg = torch.empty((n, 8), dtype=torch.int32).pin_memory(); hc = torch.empty((n, 2), dtype=torch.int32).pin_memory()
    hs = torch.empty((n, 16), dtype=torch.int32).pin_memory(); hb = torch.zeros(8, dtype=torch.int32).pin_memory()
    st = torch.cuda.current_stream().cuda_stream

    def pack():
        rc = lib.dvs_viewer_pack(*[t[k].data_ptr() for k in u.KEYS], n, g.data_ptr(), c.data_ptr(), sh.data_ptr(), bb.data_ptr(), st)
        assert rc == 0, rc

    def e2e():
        pack()
        hg.copy_(g, non_blocking=True); hc.copy_(c, non_blocking=True); hs.copy_(sh, non_blocking=True); hb.copy_(bb, non_blocking=True)

    def timed(fn):
        for _ in range(max(a.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    ms = timed(pack)
    ms_e2e = timed(e2e)
    # correctness of what was timed: the device bytes are the reference's
    got = (g.cpu().numpy().view(np.uint32), c.cpu().numpy().view(np.uint32), sh.cpu().numpy().view(np.uint32))
    exp = u.pack_with_host_ops(u.host_ops(), m)
    assert all(x.tobytes() == y.tobytes() for x, y in zip(got, exp[:3])), "timed kernel produced wrong bytes"
    peak, src = _peaks()
    ach = BYTES_PER_GAUSSIAN * n / (ms * 1e-3) / 1e9
    out = {**base, "value": n / (ms * 1e-3), "ms_per_step": ms, "scaling": "weak", "gpu_launches": 2 * a.steps,
           "roofline": {"bound": "hbm", "kernel": "viewer_pack_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": None, "peak_source": src, "alg_bytes_per_launch": BYTES_PER_GAUSSIAN * n},
           "e2e": {"value": n / (ms_e2e * 1e-3), "unit": "Gaussians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 104 * n + 32,
                   "ms_per_step": ms_e2e, "api": "dvs_viewer_pack + D2H of the three record buffers into pinned memory (what requestViewerPack queues)"},
           "cpu_baseline": cpu_reference(min(n, 1000000))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
