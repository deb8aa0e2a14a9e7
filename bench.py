#!/usr/bin/env python
"""bench.py — forward+backward Gaussians/s of the rasterize hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU; torchrun for N>1)
  python bench.py --impl reference --gpus N ...          the CPU baseline arm (oracle port on the host cores)

A "step" is one forward+backward rasterize of one view per rank (weak scaling: every extra GPU renders
one more view of the same 1M-Gaussian scene, then a single NCCL all-reduce sums the dense per-Gaussian
gradient arena).  N=1 is BASELINE config c3 (1M Gaussians, 1600x1000, SH degree 3); N>1 are views of c4.
Timed with CUDA events over exactly K steps between barrier+synchronize pairs, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "c3"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                r = [x.strip() for x in r]
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            sm.sort()
            load = [s for s in sm if s >= 0.5 * max(sm)]
            out.update(sm_mhz=load[len(load) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# Roofline side-fields (DRAM traffic per launch = dram__bytes_read.sum + dram__bytes_write.sum, issue-active %, shared-memory
# wavefront %) come from the committed summary of the `ncu --set full --clock-control none` capture of the CURRENT build:
# profiles/ncu_step.json, written by tools/ncu_summary.py from the .ncu-rep (never typed in by hand).  Valid for workload c3.
STAGE_KERNELS = {"preprocess_fwd": ("preprocess_fwd_kernel",),
                 "binning_sort": ("tile_scan_kernel", "emit_kernel", "tile_bucket_sort_kernel", "tile_bitonic_sort_kernel"),
                 "render_fwd": ("render_fwd_kernel",), "render_bwd": ("render_bwd_kernel",),
                 "preprocess_bwd": ("preprocess_bwd_kernel",)}


def ncu_side_fields(stage):
    """(traffic bytes per step, issue-active % and shared wavefront % of the stage's longest kernel, source) or Nones."""
    p = os.path.join(ROOT, "profiles", "ncu_step.json")
    if not os.path.exists(p):
        return None, None, None, None
    d = json.load(open(p))
    ks = {k: v for k, v in d["kernels"].items() if k.split("<")[0] in STAGE_KERNELS[stage]}
    if not ks:
        return None, None, None, None
    top = max(ks.values(), key=lambda v: v["time_us"])
    return (sum(v["dram_MB"] for v in ks.values()) * 1e6, top["issue_active_pct"], top["smem_wavefront_pct"],
            f"profiles/{d.get('summary', 'ncu_step.json')} (from {d['source']}, ncu --set full --clock-control none, per launch)")


def algorithmic_bytes(N, K, V, D, T, P):
    """SURVEY.md §8(d) contract figure (compulsory traffic, infinite-L2 model)."""
    pb = 44 + 12 * K
    stages = {
        "preprocess_fwd": N * pb + 48 * V,
        "binning_sort": 28 * D + 8 * T,          # tile_scan + emit + tile_sort
        "render_fwd": 4 * D + 36 * V + 20 * P,
        "render_bwd": 20 * P + 4 * D + 36 * V + 36 * V,
        "preprocess_bwd": 36 * V + V * pb + N * pb,
    }
    return stages, sum(stages.values())


_BEST_THREADS = None


def run_cpu_sample(threads=0, reps=1, frac_lin=1):
    """Oracle (port) fwd+bwd on the workload itself (frac_lin = 1: the same configuration as the GPU arm; a density-preserving
    1/frac_lin^2 crop is kept for exploration); returns Gaussians/s.
    threads <= 0: all the host threads it can use — the count is chosen once by timing the sample at the affinity
    count and at 1/2, 1/4, 1/8 of it (the OpenMP oracle stops scaling on large multi-socket hosts), best kept."""
    global _BEST_THREADS
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from divshot_b200.scenes import crop_of
    from oracle import oracle as orc
    sc = crop_of(WORKLOAD, frac_lin)
    cam = sc.cameras[0]
    oc = orc.make_camera(cam.view, cam.proj, cam.campos, cam.tanfovx, cam.tanfovy, cam.width, cam.height, cam.bg,
                         1.0, sc.sh_degree)
    arrays = (sc.means3D, sc.log_scales, sc.quats, sc.logit_opac, sc.sh0, sc.shN)

    def once(th):
        orc.set_threads(th)
        t0 = time.perf_counter()
        f = orc.forward(oc, *arrays, threads=th)
        orc.backward(oc, f, *arrays, sc.dL_dpix[0], threads=th)
        return time.perf_counter() - t0

    if threads <= 0:
        if _BEST_THREADS is None:
            # torchrun exports OMP_NUM_THREADS=1 and a container may be pinned to a subset of the host's cores:
            # start from the affinity mask
            aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            cands = sorted({max(1, aff // d) for d in (1, 2, 4, 8)}, reverse=True)
            once(cands[0])  # warm the caches / page in the library
            timing = {th: min(once(th), once(th)) for th in cands}
            _BEST_THREADS = min(timing, key=timing.get)
        threads = _BEST_THREADS
    times = [once(threads) for _ in range(reps)]
    what = f"the full workload {WORKLOAD}" if frac_lin == 1 else f"density-preserving 1/{frac_lin * frac_lin} crop of {WORKLOAD}"
    sample = (f"oracle port, {what}: {sc.N} Gaussians, "
              f"{cam.width}x{cam.height}, SH deg {sc.sh_degree}, fwd+bwd, OpenMP over Gaussians/tiles, "
              f"{threads} threads (best of the affinity count and its 1/2, 1/4, 1/8)")
    return sc.N, times, threads, sample


def choose_exchange(args, reducer, fx, grads, params, campos, deg, dev, dist, torch, fused=None, rerun_backward=None):
    """N > 1: which gradient exchange the timed steps use.  Candidates next to the reducer's plain all-reduce (NCCL / our NVLS
    kernel): `factored` (dp.FactoredGradientExchange: all-gather dL/dsh0, all-reduce 56 B/Gaussian with NCCL, dL/dshN formed by
    our kernel) and `fused` (dp.FusedGradientExchange: the same exchange as ONE kernel of ours over NVSwitch multicast, device-side
    barriers, and the backward stops writing the per-view dL/dshN).  `--allreduce factored|fused` forces one; `auto` adopts a
    candidate only if (1) on the real gradients of the warm-up step it reproduces the plain NCCL all-reduce of the whole arena
    (every tensor within 1e-4, checked on every rank, decision taken collectively so ranks can never disagree; the fused one also
    in the mode the timed steps use: gradients of a backward that skipped dL/dshN) and (2) it is the fastest here (3 timed runs
    each, max over ranks).  The returned callable carries `.bwd_flags` for the backward of the timed steps."""
    names = ("means3D", "scales", "quats", "opacities", "sh0", "shN")

    def plain():
        return reducer.all_reduce()
    plain.bwd_flags = 0
    cands = []
    if fx is not None and args.allreduce in ("auto", "factored"):
        f1 = lambda: fx.exchange(params["means3D"], campos, deg)  # noqa: E731
        f1.bwd_flags = 0
        cands.append(("factored", f1, fx))
    if fused is not None and args.allreduce in ("auto", "fused"):
        from divshot_b200 import _cabi
        f2 = lambda: fused.exchange(params["means3D"], campos, deg)  # noqa: E731
        f2.bwd_flags = _cabi.FLAG_SKIP_SHN_GRAD if rerun_backward is not None else 0
        cands.append(("fused", f2, fused))
    if not cands:
        return plain

    def timed_ms(fn, reps=3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    local = reducer.flat.clone()  # this rank's own gradients: every candidate starts from them
    ref = local.clone()
    dist.all_reduce(ref)  # plain NCCL sum of the whole arena, the semantics to reproduce
    ref_g = type(grads).allocate(grads.opacities.shape[0], grads.shN.shape[1], dev, flat=ref)

    def worst_err():
        torch.cuda.synchronize()
        err = 0.0
        for n in names:
            a, b = getattr(grads, n).double(), getattr(ref_g, n).double()
            if b.numel():
                err = max(err, float((a - b).norm() / (b.norm() + 1e-30)))
        return err

    world = dist.get_world_size()
    passed = {}
    for name, fn, obj in cands:
        ok, err = 1, float("nan")
        try:
            reducer.flat.copy_(local)
            fn()
            err = worst_err()
            if err < 1e-4 and fn.bwd_flags and rerun_backward is not None:
                rerun_backward(fn.bwd_flags)  # the mode of the timed steps: the backward no longer writes dL/dshN
                fn()
                err = max(err, worst_err())
            ok = int(err < 1e-4)
            if ok and hasattr(obj, "status") and obj.status():
                ok = 0
                sys.stderr.write(f"{name} exchange: a device-side barrier timed out\n")
        except Exception as e:  # noqa: BLE001  (a local failure must not leave the ranks disagreeing: it becomes a vote)
            ok = 0
            sys.stderr.write(f"{name} exchange unavailable: {type(e).__name__}: {e}\n")
        vote = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN)
        if int(vote.item()) == 0:
            if args.allreduce == name:
                raise RuntimeError(f"--allreduce {name}: the {name} exchange does not reproduce the plain all-reduce (rel err {err:.2e})")
            reducer.note += f"; {name} exchange rejected by its self-check (rel err {err:.2e})"
            continue
        wire = obj.wire_bytes_per_gaussian(world) if name == "factored" else obj.wire_bytes_per_gaussian()
        passed[name] = (timed_ms(fn), err, fn, wire)
    if rerun_backward is not None:
        rerun_backward(0)  # leave a complete set of local gradients behind
    else:
        reducer.flat.copy_(local)
    if not passed:
        return plain
    t_p = timed_ms(plain)
    plain_wire = 2 * (world - 1) / world * (56 + 12 * grads.shN.shape[1])
    detail = "; ".join(f"{n} exchange {t:.3f} ms (self-check rel err {e:.1e}, {w:.0f} B/Gaussian received)" for n, (t, e, _, w) in passed.items())
    detail += f" vs {reducer.backend} all-reduce {t_p:.3f} ms ({plain_wire:.0f} B/Gaussian)"
    forced = args.allreduce if args.allreduce in passed else None
    best = forced or min(passed, key=lambda n: passed[n][0])
    if forced or passed[best][0] < t_p:
        reducer.note = (reducer.note + "; " if reducer.note else "") + detail
        reducer.backend = best
        return passed[best][2]
    reducer.note = (reducer.note + "; " if reducer.note else "") + "kept: " + detail
    return plain


def measure_row(tool):
    """Rows F3 (trainer -> viewer hand-off) and F1 (refinement step) of SURVEY.md §8 f, each measured by its own harness
    (tools/bench_viewer_pack.py, tools/bench_densify.py) in a SUBPROCESS after the headline measurement is complete: the
    rows were built when round 1 had no GPU minutes left, so this is their first run on a B200 — a failure there must not
    be able to disturb the headline number (separate CUDA context, time-boxed), and is reported as text instead."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), "--steps", "30", "--warmup", "5"],
                           capture_output=True, text=True, timeout=300, cwd=ROOT)
        rows = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not rows:
            return {"error": (r.stderr or r.stdout)[-600:]}
        d = json.loads(rows[-1])
        return {k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "config", "roofline", "e2e", "cpu_baseline", "gpu_launches",
                                      "passes", "refine_ms_per_call", "refine_ms_per_iteration", "model_2dgs") if k in d}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:600]}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        run_cpu_sample(reps=1)
    n, times, cores, sample = run_cpu_sample(reps=args.steps)
    from divshot_b200.scenes import CONFIGS
    _, _, W, H, deg, _ = CONFIGS[WORKLOAD]
    total = sum(times)
    val = n * args.steps / total
    line = {"impl": "reference", "metric": "fwd+bwd Gaussians/s", "value": val, "unit": "Gaussians/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same workload as the GPU arm at N=1, whole (not a crop), on the host cores
            "config": {"workload": f"{WORKLOAD}: {n} Gaussians, {W}x{H}, SH deg {deg}, 1 view per rank per step",
                       "N": n, "width": W, "height": H, "sh_degree": deg, "views_per_step": 1},
            "cpu_baseline": {"value": val, "unit": "Gaussians/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Gaussians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference ships no implementation of this path (SURVEY.md §0): kind=port is the CPU oracle"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="override: c2|c3|c5 (parity/bench exploration only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-rows", action="store_true", help="skip the sub-process measurement of the other section-8 rows")
    ap.add_argument("--no-bind", action="store_true",
                    help="do not pin the process to the CPUs of its GPU's NUMA node (default: pinned, so that the pinned host "
                         "buffers of the end-to-end path are node-local; the CPU baseline leg runs with the original affinity)")
    ap.add_argument("--no-tight", action="store_true",
                    help="timed steps keep the whole-rectangle tile lists (default: DVS_FLAG_TIGHT_LISTS — entries whose sub-tile "
                         "mask is empty are not emitted; image and gradients are bit-identical, tests/test_gpu_parity.py)")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nvls", "nccl", "factored", "fused"],
                    help="gradient exchange: NVSwitch in-switch reduction over symmetric memory, plain NCCL, the factored "
                         "exchange (all-gather dL/dsh0, all-reduce 56 B/Gaussian, form dL/dshN locally: 3 NCCL calls + our kernel) "
                         "or the same exchange fused into ONE kernel of ours over NVSwitch multicast; auto = the fastest of them, "
                         "the factored / fused ones only after they have reproduced the plain all-reduce on this run's gradients")
    ap.add_argument("--fused-ctas", type=int, default=0, help="fused exchange: CTAs of the kernel (0 = one per SM)")
    ap.add_argument("--fused-reduce-ctas", type=int, default=0, help="fused exchange: CTAs that do the in-switch reduction (0 = half)")
    args = ap.parse_args()
    global WORKLOAD
    if args.workload:
        WORKLOAD = args.workload
    if args.impl == "reference":
        return reference_arm(args)

    # only the JSON line may reach stdout (NCCL prints its version banner there when NCCL_DEBUG is set)
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from divshot_b200 import _cabi
    from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
    from divshot_b200.dp import GradientReducer
    from divshot_b200.scenes import CONFIGS, make_scene

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    bind_note = "not bound (--no-bind)"
    if not args.no_bind:
        from divshot_b200.hostbind import bind_to_gpu_numa_node
        bind_note = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # scene: same Gaussians on every rank (replicas), one distinct view per rank
    views = max(world, 1)
    sc = make_scene(WORKLOAD, views=views) if world > 1 else make_scene(WORKLOAD)
    _, N, W, H, deg, _ = CONFIGS[WORKLOAD]
    K = (deg + 1) ** 2
    cam = _cabi.make_camera(sc.cameras[rank % len(sc.cameras)], deg)
    params = scene_to_device(sc, dev)
    dl_host = torch.from_numpy(sc.dL_dpix[rank % len(sc.dL_dpix)]).pin_memory()
    dl = dl_host.to(dev)
    img_host = torch.empty(3, H, W, dtype=torch.float32).pin_memory()
    factored = args.allreduce == "factored"
    reducer = GradientReducer(GradBuffers.numel_for(N, K - 1), dev,
                              backend="nccl" if factored else "auto" if args.allreduce == "fused" else args.allreduce)
    grads = GradBuffers.allocate(N, K - 1, dev, flat=reducer.flat)
    exchange = reducer.all_reduce
    fx = fused = None
    campos = torch.tensor(np.asarray(sc.cameras[rank % len(sc.cameras)].campos, np.float32))
    if world > 1 and args.allreduce in ("auto", "fused") and K > 1:
        ok = 1
        try:
            from divshot_b200.dp import FusedGradientExchange
            fused = FusedGradientExchange(grads, reducer, ctas=args.fused_ctas, reduce_ctas=args.fused_reduce_ctas)
            fused.set_cameras(campos)
        except Exception as e:  # noqa: BLE001
            ok, fused = 0, None
            if args.allreduce == "fused":
                raise
            sys.stderr.write(f"fused exchange not set up: {type(e).__name__}: {e}\n")
        vote = torch.tensor([ok], device=dev, dtype=torch.int32)  # all ranks use it or none does
        dist.all_reduce(vote, op=dist.ReduceOp.MIN)
        if int(vote.item()) == 0:
            fused = None
    if world > 1 and args.allreduce in ("auto", "factored"):
        try:
            from divshot_b200.dp import FactoredGradientExchange
            fx = FactoredGradientExchange(grads)
            fx.set_cameras(campos)
        except Exception as e:  # noqa: BLE001  (same code and arguments on every rank: fails on all of them or on none)
            if args.allreduce == "factored":
                raise
            sys.stderr.write(f"factored exchange not set up: {type(e).__name__}: {e}\n")
            fx = None
    rast = Rasterizer(local)
    rast.reserve(N, W, H, 0)
    img = torch.empty(3, H, W, device=dev)
    radii = torch.empty(N, dtype=torch.int32, device=dev)

    # the arena has been sized by the synchronous warm-up forwards below; timed steps run without any per-step
    # host synchronisation (DVS_FLAG_DEFER_CHECK) — an overflow would surface as an error at rast.stats()
    rast.forward(cam, params, img, radii); rast.backward(dl, grads)
    rast.forward(cam, params, img, radii); rast.backward(dl, grads)

    def rerun_backward(flags):
        rast.forward(cam, params, img, radii)
        rast.backward(dl, grads, flags=flags)

    bwd_flags = 0
    if fx is not None or fused is not None:
        exchange = choose_exchange(args, reducer, fx, grads, params, campos, deg, dev, dist, torch, fused=fused,
                                   rerun_backward=rerun_backward)
        bwd_flags = getattr(exchange, "bwd_flags", 0)

    # the training-loop mode: no host synchronisation per step, single-pass binning, tight tile lists
    cam_defer = _cabi.DvsCamera.from_buffer_copy(cam)
    cam_defer.flags |= _cabi.FLAG_DEFER_CHECK | (0 if args.no_tight else _cabi.FLAG_TIGHT_LISTS)

    def step_resident():
        rast.forward(cam_defer, params, img, radii, defer_check=True)
        rast.backward(dl, grads, flags=bwd_flags)
        if world > 1:
            exchange()

    # end to end through the C-ABI with HOST image buffers, as a pipelined trainer drives it (dvs_rast_step_host_async / _wait):
    # two slots of pinned buffers; step k is queued (H2D of its dL/dpix, forward, D2H of its image, backward, the exchange) before
    # the host waits for step k-1's image, so the copies of neighbouring steps overlap compute and the launch queue never drains.
    # Every step's H2D and D2H are inside the timed region; drain() waits for the last image.
    dl_hosts = [dl_host, dl_host.clone().pin_memory()]
    img_hosts = [img_host, torch.empty_like(img_host).pin_memory()]
    e2e_state = {"k": 0}

    def step_e2e():
        k = e2e_state["k"]
        rast.step_host_async(cam_defer, params, grads, dl_hosts[k & 1], img_hosts[k & 1], k & 1, flags=bwd_flags)
        if world > 1:
            exchange()
        if k > 0:
            rast.step_host_wait((k - 1) & 1)
        e2e_state["k"] = k + 1

    def drain_e2e():
        if e2e_state["k"] > 0:
            rast.step_host_wait((e2e_state["k"] - 1) & 1)
    step_e2e.drain = drain_e2e

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if sample_clocks else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if hasattr(fn, "drain"):
            fn.drain()  # the pipelined end-to-end loop: the last step's image has arrived on the host
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    warm = max(args.warmup, 3)
    sampler_all = ClockSampler(local)  # nvidia-smi sampling across warm-up + all timed regions (>= a few samples)
    # the library's per-stage CUDA events (nine records per step) are a measuring device, not part of the step: off in the
    # timed loops as in a training loop, on for the per-stage breakdown below
    rast.set_profiling(False)
    ms_total, _ = timed(step_resident, args.steps, warm)
    # kernels of the library per step: the context's own launch counter (dvs_rast_kernel_launches) around one more step
    try:
        l0 = rast.kernel_launches(); step_resident(); launches_per_step = rast.kernel_launches() - l0
    except Exception:  # an older library without the counter: preprocess, scan, 3 sort classes, 2 compositing, per-Gaussian backward
        launches_per_step = 8
    rast.set_profiling(True)
    # per-stage device times (CUDA events recorded on the launch stream inside the library), averaged over a few steps
    stage_ms = {}
    reps = 5
    for _ in range(reps):
        rast.forward(cam_defer, params, img, radii, defer_check=True)  # same mode as the timed steps
        rast.backward(dl, grads)
        for k, v in rast.stage_ms().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v / reps
    st = rast.stats()
    rast.set_profiling(False)
    ms_e2e, _ = timed(step_e2e, args.steps, warm)
    # keep the GPU under the same load a little longer so the 100 ms nvidia-smi sampler sees it
    t_end = time.time() + 0.6
    while time.time() < t_end:
        step_resident()
    torch.cuda.synchronize()
    clocks = sampler_all.stop()
    ms_ar = None
    if world > 1:  # the collective alone (device time, max over ranks), for the scaling breakdown
        ms_ar, _ = timed(lambda: exchange(), args.steps, warm)
        ms_ar /= args.steps

    ms_step = ms_total / args.steps
    value = N * world / (ms_step * 1e-3)
    e2e_value = N * world / (ms_e2e / args.steps * 1e-3)
    peak, peak_src = _peaks()
    V, D, T, P = st["num_visible"], st["num_dups"], st["tiles_x"] * st["tiles_y"], W * H
    stage_bytes, total_bytes = algorithmic_bytes(N, K, V, D, T, P)
    merged = {"preprocess_fwd": stage_ms["preprocess_fwd"],
              "binning_sort": stage_ms["tile_scan"] + stage_ms["emit"] + stage_ms["tile_sort"],
              "render_fwd": stage_ms["render_fwd"], "render_bwd": stage_ms["render_bwd"],
              "preprocess_bwd": stage_ms["preprocess_bwd"]}
    dominant = max(merged, key=merged.get)
    stages = {k: {"ms": round(merged[k], 4), "alg_MB": round(stage_bytes[k] / 1e6, 2),
                  "GBps": round(stage_bytes[k] / 1e9 / (merged[k] * 1e-3), 1) if merged[k] > 0 else None,
                  "frac_hbm": round(stage_bytes[k] / 1e9 / (merged[k] * 1e-3) / peak, 4) if merged[k] > 0 else None}
              for k in merged}
    dom_gbs = stage_bytes[dominant] / 1e9 / (merged[dominant] * 1e-3)
    side = ncu_side_fields(dominant) if WORKLOAD == "c3" else (None, None, None, None)
    for k in stages:  # DRAM traffic per stage from the same capture (well above alg_MB = wasted re-reads)
        stages[k]["ncu_dram_MB"] = (round(ncu_side_fields(k)[0] / 1e6, 2) if WORKLOAD == "c3" and ncu_side_fields(k)[0] else None)
    step_gbs = total_bytes / 1e9 / (ms_step * 1e-3) if world == 1 else None

    line = {
        "metric": "fwd+bwd Gaussians/s", "value": value, "unit": "Gaussians/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: {N} Gaussians, {W}x{H}, SH deg {deg}, 1 view per rank per step"
                               + (", NCCL all-reduce of the dense gradient arena" if world > 1 else ""),
                   "N": N, "width": W, "height": H, "sh_degree": deg, "views_per_step": world,
                   "visible": V, "duplicates": D, "tiles": T, "max_tile_len": st["max_tile_len"],
                   "list_entries": st["num_list_entries"],
                   "tile_lists": ("whole-rectangle (reference-exact)" if args.no_tight else
                                  "tight: the reference-exact lists minus the entries whose sub-tile mask is empty "
                                  "(DVS_FLAG_TIGHT_LISTS; image and gradients bit-identical, D counts all duplicates)"),
                   "l2": "inputs larger than L2 (params+grads 472 MB + 96 MB records/lists per step); no explicit flush",
                   "parallelism": f"dp{world} (view-sharded replicas)", "host_binding": bind_note,
                   "host_sync": "none per step (binning arena validated by deferred check, DVS_FLAG_DEFER_CHECK; "
                                "single-pass binning into fixed-stride tile bins sized by the warm-up forwards)"},
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
                     "frac": dom_gbs / peak,
                     "traffic": side[0], "traffic_source": side[3], "issue_active_pct": side[1], "smem_wavefront_pct": side[2],
                     "peak_source": peak_src,
                     "note": "algorithmic bytes per SURVEY.md §8(d) / CUDA-event stage time.  The dominant kernel is the "
                             "compositing backward: 256*D potential pair evaluations against ~0.12 GB of compulsory "
                             "traffic, so it is bound by instruction issue and shared-memory wavefronts (see issue_active_pct, "
                             "smem_wavefront_pct), not by HBM; the "
                             "streaming stages' HBM fractions are in `stages`"},
        "roofline_step": ({"achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                           "alg_bytes_per_step": total_bytes} if step_gbs else None),
        "stages": stages,
        "e2e": {"value": e2e_value, "unit": "Gaussians/s", "h2d_bytes_per_step": 12 * P * world,
                "d2h_bytes_per_step": 12 * P * world, "ms_per_step": ms_e2e / args.steps,
                "api": "dvs_rast_step_host_async / dvs_rast_step_host_wait (C-ABI), two pipeline slots: every step copies its pinned "
                       "dL/dpix H2D and its image D2H (the host waits for step k's image after queueing step k+1); parameters and "
                       "gradients device-resident as in the trainer"},
        # kernels of ours in the timed region: the rasterizer's own count per step (6 forward + 2 backward at c3); with N > 1
        # one more when the exchange is ours too (the fused exchange / NVLS all-reduce kernel, or the SH accumulation kernel
        # of the factored exchange)
        "gpu_launches": (launches_per_step + (1 if world > 1 and reducer.backend in ("nvls", "factored", "fused") else 0)) * args.steps,
        "allreduce": ({"backend": reducer.backend, "note": reducer.note, "bytes": int(reducer.flat.numel()) * 4, "ms": ms_ar,
                       "busbw_GBps": (2 * (world - 1) / world * reducer.flat.numel() * 4 / 1e9 / (ms_ar * 1e-3))}
                      if world > 1 else None),
        "clocks": clocks,
    }
    if affinity0 is not None:
        os.sched_setaffinity(0, affinity0)  # the CPU legs below use every core the process was given
    if world == 1 and rank == 0 and not args.no_cpu:
        n_s, times, cores, sample = run_cpu_sample(reps=3)
        best = min(times)
        line["cpu_baseline"] = {"value": n_s / best, "unit": "Gaussians/s", "cores": cores, "kind": "port",
                                "sample": sample + f"; best of 3 ({best:.2f} s)"}
    if world == 1 and rank == 0 and not args.no_rows:
        line["other_rows"] = {"F3_viewer_pack": measure_row("bench_viewer_pack.py"), "F1_refinement": measure_row("bench_densify.py"),
                              "F1_train_step": measure_row("bench_trainstep.py")}
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    rast.close()


if __name__ == "__main__":
    main()
