#!/usr/bin/env python
"""A/B harness for kernel experiments: several builds / runtime modes of the rasterizer, one GPU call.

  python tools/ab_bench.py --variants r1 default default+tight v_x+tight --workload c3 --steps 30 --out gpurun_out/ab.json

A variant is `<lib>[+tight][+twopass][+absgrad][+2dgs]`: <lib> = "default" (divshot_b200/lib/libdvsrast.so) or the name of a
`divshot_b200.build.build_variant` build (divshot_b200/lib/variants/libdvsrast_<lib>.so).  Each variant runs in its
own process (fresh CUDA context, DVS_RAST_LIB), renders the workload, and is compared with the FIRST variant's outputs:
image bit-identical?, max abs image difference, n_contrib equal?, norm-wise relative error of every gradient tensor.
Timing: CUDA events over `--steps` device-resident steps in the training-loop mode (deferred check), plus the library's
own per-stage events averaged over 10 steps.  Not a bench line (bench.py is); this decides what goes INTO the build.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lib_path(lib):
    if lib == "default":
        return os.path.join(ROOT, "divshot_b200", "lib", "libdvsrast.so")
    return os.path.join(ROOT, "divshot_b200", "lib", "variants", f"libdvsrast_{lib}.so")


def child(args):
    import numpy as np
    import torch
    from divshot_b200 import _cabi
    from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
    from divshot_b200.scenes import CONFIGS, make_scene

    spec = args.child.split("+")
    tight, twopass = "tight" in spec[1:], "twopass" in spec[1:]
    absgrad = "absgrad" in spec[1:]
    surfel = "2dgs" in spec[1:]  # the 2DGS variant (DVS_FLAG_MODEL_2DGS): its stage times; the comparison columns are meaningless
    _, N, W, H, deg, _ = CONFIGS[args.workload]
    K = (deg + 1) ** 2
    sc = make_scene(args.workload)
    dev = torch.device("cuda", 0)
    params = scene_to_device(sc, dev)
    flags = 0
    if tight:
        flags |= getattr(_cabi, "FLAG_TIGHT_LISTS", 32)
    if surfel:
        flags |= getattr(_cabi, "FLAG_MODEL_2DGS", 128)
    cam = _cabi.make_camera(sc.cameras[0], deg, flags=flags)
    dl = torch.from_numpy(sc.dL_dpix[0]).to(dev)
    grads = GradBuffers.allocate(N, K - 1, dev)
    rast = Rasterizer(0)
    img = torch.empty(3, H, W, device=dev)
    radii = torch.empty(N, dtype=torch.int32, device=dev)
    m2a = torch.zeros(N, 2, device=dev) if absgrad else None
    rast.forward(cam, params, img, radii); rast.backward(dl, grads, mean2D_abs=m2a)
    rast.forward(cam, params, img, radii); rast.backward(dl, grads, mean2D_abs=m2a)
    defer = not twopass

    def step():
        rast.forward(cam, params, img, radii, defer_check=defer)
        rast.backward(dl, grads, mean2D_abs=m2a)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    rast.set_profiling(False)  # the timed loop runs as a training loop does: no per-stage events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / args.steps)
    rast.set_profiling(True)
    stage = {}
    for _ in range(10):
        step()
        for k, v in rast.stage_ms().items():
            stage[k] = stage.get(k, 0.0) + v / 10
    st = rast.stats()
    out = {"variant": args.child, "ms_step": round(best, 4), "stages": {k: round(v, 4) for k, v in stage.items() if v},
           "D": st["num_dups"], "list_entries": st.get("num_list_entries", st["num_dups"]), "max_tile_len": st["max_tile_len"]}
    # outputs of the timed mode, compared with the first variant's
    step(); torch.cuda.synchronize()
    cur = {"image": img.cpu().numpy(), "n_contrib": rast.debug_read(_cabi.BUF_N_CONTRIB),
           "final_T": rast.debug_read(_cabi.BUF_FINAL_T)}
    for n in ("means3D", "scales", "quats", "opacities", "sh0", "shN"):
        cur["g_" + n] = getattr(grads, n).cpu().numpy()
    if absgrad:
        cur["g_mean2D_abs"] = m2a.cpu().numpy()
    if not os.path.exists(args.ref):
        np.savez(args.ref, **cur)
        out["cmp"] = "reference variant"
    else:
        ref = np.load(args.ref)
        cmp = {"image_bit_identical": bool(np.array_equal(cur["image"], ref["image"])),
               "image_max_abs": float(np.abs(cur["image"] - ref["image"]).max()),
               "final_T_bit_identical": bool(np.array_equal(cur["final_T"], ref["final_T"])),
               "n_contrib_equal": bool(np.array_equal(cur["n_contrib"], ref["n_contrib"]))}
        for k in cur:
            if k.startswith("g_") and k in ref:
                a, b = cur[k].astype(np.float64), ref[k].astype(np.float64)
                cmp[k] = float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
        out["cmp"] = cmp
    print("AB " + json.dumps(out), flush=True)
    rast.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", nargs="+", default=["default"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--out", default=None)
    ap.add_argument("--child", default=None)
    ap.add_argument("--ref", default="/tmp/ab_ref.npz")
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.child:
        return child(args)
    if os.path.exists(args.ref):
        os.unlink(args.ref)
    rows = []
    for v in args.variants:
        lib = v.split("+")[0]
        env = dict(os.environ, DVS_RAST_LIB=lib_path(lib))
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", v, "--workload", args.workload, "--steps",
                                str(args.steps), "--ref", args.ref], env=env, capture_output=True, text=True, timeout=args.timeout)
            line = [l for l in r.stdout.splitlines() if l.startswith("AB ")]
            row = json.loads(line[-1][3:]) if line else {"variant": v, "error": (r.stderr or r.stdout)[-800:]}
        except subprocess.TimeoutExpired:
            row = {"variant": v, "error": "timeout"}
        row["wall_s"] = round(time.time() - t0, 1)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
