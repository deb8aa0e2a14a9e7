#!/usr/bin/env python
"""Measurement for row F1 (the trainer's refinement step, csrc/densify.cu), same conventions as bench.py: one JSON line.

  python tools/bench_densify.py [--n 1000000] [--steps 30] [--warmup 5]

Per-iteration passes (run every step inside the refinement window), each an HBM-bound stream over N Gaussians:
  mcmc_noise       reads scales 12 + quats 16 + opacity 4 + means 12, writes means 12            = 56 B/Gaussian
  mcmc_regularise  reads opacity 4 + scales 12, read-modify-writes grad opacity 8 + scales 24   = 48 B/Gaussian
  adc_accumulate   reads mean2D grad 8 + radii 4, read-modify-writes accum 8 + denom 8          = 28 B/Gaussian
`value` = Gaussians/s of the MCMC per-iteration work (noise + regularise); roofline = their algorithmic bytes / event time
vs the measured HBM peak.  The every-`refineEvery` step itself (relocation + 5 % growth: scan, sampling, row copies; queued
without a host synchronisation, persistent workspace) is reported as device milliseconds per call, amortised over
refineEvery = 100 in `refine_ms_per_iteration`.
No CPU baseline exists: the closed trainer's implementation is absent from the reference."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import densify_ref as dr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    import torch
    from bench import _peaks
    from divshot_b200 import build
    assert torch.cuda.is_available(), "needs a GPU (no CPU path in the product)"
    torch.zeros(1, device="cuda")
    lib = C.CDLL(build.build_gstrain())
    vp, ll, f, ull = C.c_void_p, C.c_longlong, C.c_float, C.c_ulonglong
    lib.dvs_densify_test_mcmc_noise.argtypes = [vp, ll, f, ull, vp]
    lib.dvs_densify_test_mcmc_regularise.argtypes = [vp, vp, ll, f, f, vp]
    lib.dvs_densify_test_adc_accumulate.argtypes = [vp] * 5 + [ll, vp]
    lib.dvs_densify_test_mcmc_refine.argtypes = [vp, vp, vp, C.POINTER(ll), ll, ll, f, ull, C.POINTER(ll), vp]
    N, cap = a.n, int(a.n * 1.1)
    m = dr.random_model(N, cap, 1)
    P = {k: torch.from_numpy(m[k]).cuda() for k in dr.KEYS}
    G = {k: torch.zeros_like(v) for k, v in P.items()}
    M1 = {k: torch.zeros_like(v) for k, v in P.items()}
    M2 = {k: torch.zeros_like(v) for k, v in P.items()}
    tab = lambda d: (C.c_void_p * 6)(*[d[k].data_ptr() for k in dr.KEYS])  # noqa: E731
    tp, tg, t1, t2 = tab(P), tab(G), tab(M1), tab(M2)
    g2 = torch.randn(cap, 2, device="cuda"); radii = torch.randint(-1, 5, (cap,), dtype=torch.int32, device="cuda")
    acc = torch.zeros(cap, device="cuda"); den = torch.zeros(cap, device="cuda")

    def timed(fn, steps=a.steps):
        for _ in range(max(a.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def ok(rc):
        assert rc == 0, rc

    ms_noise = timed(lambda: ok(lib.dvs_densify_test_mcmc_noise(tp, N, 1e-6, 7, None)))
    ms_reg = timed(lambda: ok(lib.dvs_densify_test_mcmc_regularise(tp, tg, N, 0.01, 0.01, None)))
    ms_adc = timed(lambda: ok(lib.dvs_densify_test_adc_accumulate(g2.data_ptr(), None, radii.data_ptr(), acc.data_ptr(), den.data_ptr(), N, None)))

    def refine():
        n = ll(N)  # no report requested: the call only queues work (the dead count stays on the device)
        ok(lib.dvs_densify_test_mcmc_refine(tp, t1, t2, C.byref(n), cap, cap, 0.005, 11, None, None))
        return n.value

    ms_refine = timed(refine, steps=max(3, a.steps // 10))
    peak, src = _peaks()
    gbs = lambda b, ms: b * N / (ms * 1e-3) / 1e9  # noqa: E731
    per_iter = ms_noise + ms_reg
    ach = gbs(56 + 48, per_iter)
    print(json.dumps({
        "metric": "refinement per-iteration passes, Gaussians/s", "value": N / (per_iter * 1e-3), "unit": "Gaussians/s", "n_gpus": 1,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": per_iter, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"F1: {N} Gaussians (arena {cap}), MCMC strategy: exploration noise + regularisers every iteration"},
        "roofline": {"bound": "hbm", "kernel": "k_noise + k_regularise", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "peak_source": src, "alg_bytes_per_launch": (56 + 48) * N},
        "passes": {"mcmc_noise": {"ms": ms_noise, "GBps": gbs(56, ms_noise)}, "mcmc_regularise": {"ms": ms_reg, "GBps": gbs(48, ms_reg)},
                   "adc_accumulate": {"ms": ms_adc, "GBps": gbs(28, ms_adc)}},
        "refine_ms_per_call": ms_refine, "refine_ms_per_iteration": ms_refine / 100.0,
        "cpu_baseline": None, "gpu_launches": 2 * a.steps}))


if __name__ == "__main__":
    main()
