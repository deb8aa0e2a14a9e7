#!/usr/bin/env python
"""Multi-GPU parity on hardware: the EXCHANGED gradient arena against the sum of the ORACLE's per-view gradients.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      tools/check_dp_vs_oracle.py --workload c3 --out gpurun_out/dp_vs_oracle_n2.json

Every rank renders its own ring view of the scene with the CUDA path (forward + backward through the C-ABI), runs the live
OpenMP oracle on the same view, and the oracle gradients are summed over ranks in float64 (one all-reduce of doubles — the
checker's sum, not the product's).  Then each exchange the bench can pick — plain NCCL, our NVLS all-reduce kernel, the factored
exchange (3 NCCL calls + our SH kernel), the fused exchange (one kernel of ours, also with the backward skipping dL/dshN) — is
run on the CUDA gradients and every tensor is compared with that sum: norm-wise relative error, share of elements within 1e-4
(tests/util.assert_close_robust's metric).  The oracle (test infrastructure) is only the checker here."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from divshot_b200 import _cabi
    from divshot_b200.dp import FactoredGradientExchange, FusedGradientExchange, GradientReducer
    from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
    from divshot_b200.scenes import CONFIGS, make_scene
    from oracle import oracle as orc
    from util import elem_err, oracle_threads, orc_cam, scene_arrays

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    _, N, W, H, deg, _ = CONFIGS[args.workload]
    K = (deg + 1) ** 2
    sc = make_scene(args.workload, views=world)
    params = scene_to_device(sc, dev)
    cam = _cabi.make_camera(sc.cameras[rank], deg)
    dl = torch.from_numpy(sc.dL_dpix[rank]).to(dev)
    reducer = GradientReducer(GradBuffers.numel_for(N, K - 1), dev, backend="auto")
    g = GradBuffers.allocate(N, K - 1, dev, flat=reducer.flat)
    rast = Rasterizer(local)

    # the checker: oracle gradients of my view, summed over ranks in float64
    th = max(1, oracle_threads() // world)
    orc.set_threads(th)
    oc = orc_cam(sc.cameras[rank], deg)
    f = orc.forward(oc, *scene_arrays(sc), threads=th)
    b = orc.backward(oc, f, *scene_arrays(sc), sc.dL_dpix[rank], threads=th)
    names = ("means3D", "scales", "quats", "opacities", "sh0", "shN")
    ref = {}
    for n, a in zip(names, (b.dL_dmeans3D, b.dL_dscales, b.dL_dquats, b.dL_dopacities, b.dL_dsh0, b.dL_dshN)):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        dist.all_reduce(t)
        ref[n] = t.cpu().numpy()

    def backward(flags=0):
        rast.forward(cam, params)
        g.flat.fill_(float("nan")) if not flags else None
        rast.backward(dl, g, flags=flags)
        torch.cuda.synchronize()

    def compare(what):
        torch.cuda.synchronize()
        row = {"exchange": what, "rank": rank}
        worst = 0.0
        for n in names:
            a = getattr(g, n).cpu().numpy().astype(np.float64).reshape(ref[n].shape)
            nrm = float(np.linalg.norm(a - ref[n]) / (np.linalg.norm(ref[n]) + 1e-30))
            e = elem_err(a, ref[n])
            row[n] = {"norm_rel_err": nrm, "within_1e-4": float((e <= 1e-4).mean()), "worst_elem": float(e.max())}
            worst = max(worst, nrm)
        row["worst_norm_rel_err"] = worst
        row["pass_1e-4"] = bool(worst <= 1e-4 and all(row[n]["within_1e-4"] >= 0.999 for n in names))
        return row

    rows = []
    campos = torch.tensor(np.asarray(sc.cameras[rank].campos, np.float32))
    backward(); dist.all_reduce(reducer.flat); rows.append(compare("plain NCCL all-reduce"))
    if reducer._hdl is not None:
        keep = reducer.backend
        reducer.backend = "nvls"
        backward(); reducer.all_reduce(); rows.append(compare("our NVLS all-reduce kernel (multimem two-shot)"))
        reducer.backend = keep
    fx = FactoredGradientExchange(g)
    fx.set_cameras(campos)
    backward(); fx.exchange(params["means3D"], campos, deg); rows.append(compare("factored exchange (3 NCCL calls + our SH kernel)"))
    if reducer._hdl is not None:
        fu = FusedGradientExchange(g, reducer)
        fu.set_cameras(campos)
        backward(); fu.exchange(params["means3D"], campos, deg); rows.append(compare("fused exchange (one kernel of ours)"))
        backward(_cabi.FLAG_SKIP_SHN_GRAD); fu.exchange(params["means3D"], campos, deg)
        rows.append(compare("fused exchange, backward with DVS_FLAG_SKIP_SHN_GRAD (the bench's mode)"))
        rows[-1]["barrier_status"] = fu.status()
    allrows = [None] * world
    dist.all_gather_object(allrows, rows)
    if rank == 0:
        out = {"workload": args.workload, "world": world, "N": N, "image": [W, H], "sh_degree": deg,
               "reference": "sum over ranks (float64) of the OpenMP oracle's per-view gradients", "rows": [r for rr in allrows for r in rr]}
        text = json.dumps(out, indent=1)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            open(args.out, "w").write(text)
        for r in out["rows"]:
            print(f"rank {r['rank']} {r['exchange']}: worst norm-wise rel err {r['worst_norm_rel_err']:.2e} pass {r['pass_1e-4']}")
    dist.destroy_process_group()
    rast.close()


if __name__ == "__main__":
    main()
