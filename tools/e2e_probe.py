#!/usr/bin/env python
"""Where the end-to-end step loses time against the device-resident one (1 GPU): the resident step timed (a) alone, (b) with a
19.2 MB H2D and a 19.2 MB D2H running concurrently on side streams each step (no dependencies), and the pipelined host step
(dvs_rast_step_host_async / _wait) (c) as bench.py drives it, (d) waiting two steps late instead of one."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from divshot_b200 import _cabi
    from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
    from divshot_b200.scenes import CONFIGS, make_scene
    _, N, W, H, deg, _ = CONFIGS["c3"]
    sc = make_scene("c3")
    dev = torch.device("cuda", 0)
    params = scene_to_device(sc, dev)
    cam = _cabi.make_camera(sc.cameras[0], deg)
    cam_d = _cabi.make_camera(sc.cameras[0], deg, flags=_cabi.FLAG_DEFER_CHECK | _cabi.FLAG_TIGHT_LISTS)
    dl_h = [torch.from_numpy(sc.dL_dpix[0]).pin_memory() for _ in range(3)]
    img_h = [torch.empty(3, H, W).pin_memory() for _ in range(3)]
    dl = dl_h[0].to(dev)
    grads = GradBuffers.allocate(N, 15, dev)
    rast = Rasterizer(0)
    img = torch.empty(3, H, W, device=dev); radii = torch.empty(N, dtype=torch.int32, device=dev)
    rast.forward(cam, params, img, radii); rast.backward(dl, grads)
    rast.forward(cam, params, img, radii); rast.backward(dl, grads)
    rast.set_profiling(False)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    side_in, side_out = torch.empty_like(dl), torch.randn(3, H, W, device=dev)

    def resident(copies):
        def f():
            if copies:
                with torch.cuda.stream(s_in):
                    side_in.copy_(dl_h[1], non_blocking=True)
                with torch.cuda.stream(s_out):
                    img_h[2].copy_(side_out, non_blocking=True)
            rast.forward(cam_d, params, img, radii, defer_check=True)
            rast.backward(dl, grads)
        return f

    def timed(fn, steps=40, warm=8, drain=None):
        for _ in range(warm):
            fn()
        if drain:
            drain()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain:
            drain()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    out = {"resident_ms": timed(resident(False)), "resident_with_concurrent_copies_ms": timed(resident(True))}
    for lag in (1, 2):
        st = {"k": 0}

        def step():
            k = st["k"]
            rast.step_host_async(cam_d, params, grads, dl_h[k & 1], img_h[k & 1], k & 1)
            if k >= lag:
                rast.step_host_wait((k - lag) & 1) if lag == 1 else None
            st["k"] = k + 1

        def drain():
            rast.step_host_wait((st["k"] - 1) & 1)
        out[f"pipelined_wait_lag{lag}_ms" if lag == 1 else "pipelined_no_wait_until_drain_ms"] = timed(step, drain=drain)
    rast.set_profiling(True)
    out["resident_profiling_events_on_ms"] = timed(resident(False))
    print(json.dumps(out))
    rast.close()


if __name__ == "__main__":
    main()
