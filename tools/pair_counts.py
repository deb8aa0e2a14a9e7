#!/usr/bin/env python
"""Secondary (honesty) counters of SURVEY.md §8(d) that need no GPU: how many (pixel, splat) pair evaluations the two
compositing kernels have to do, counted on the CPU from the ORACLE's sorted tile lists with a numpy restatement of the
kernels' culling rules (sub-tile mask of csrc/emit.cuh, warp early-out of csrc/render_fwd.cu, deepest-contributor bound
of csrc/render_bwd.cu).  Explains why those kernels are bound by issue slots, not HBM, and what each culling level buys.

    python tools/pair_counts.py [--config c2 | --crop c3:2]      (c3:2 = density-preserving 1/4 crop of c3)

Prints one JSON object; profiles/r1_pair_counts.md holds the committed runs.  Test infrastructure (imports oracle/)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LOG2E = np.float32(1.4426950408889634)
ALPHA_MIN_LOG2 = np.float32(-7.994353436858858)


def sub_tile_masks(mx, my, a, b, c, m, X0, Y0):
    """emit.cuh sub_tile_mask for arrays of entries -> uint8 masks (bit 2*ry+cx: 8x4 box cx, ry of the 16x16 tile)."""
    f = np.float32
    ok = (a > 0) & (c > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        hbc = np.where(ok, f(-0.5) * b / c, f(0))
        hba = np.where(ok, f(-0.5) * b / a, f(0))
    m = np.where(ok, m, f(3.0e38))
    rx, ry_ = mx - X0, my - Y0
    mask = np.zeros(mx.shape, np.uint8)
    for ry in range(4):
        dyhi = ry_ - f(4 * ry); dylo = dyhi - f(3)
        eyn = np.minimum(np.maximum(f(0), dylo), dyhi)
        dxs = hba * eyn
        for cx in range(2):
            dxhi = rx - f(8 * cx); dxlo = dxhi - f(7)
            ex = np.minimum(np.maximum(f(0), dxlo), dxhi)
            dy = np.minimum(np.maximum(hbc * ex, dylo), dyhi)
            q1 = a * ex * ex + (b * ex + c * dy) * dy
            dx = np.minimum(np.maximum(dxs, dxlo), dxhi)
            q2 = c * eyn * eyn + (b * eyn + a * dx) * dx
            hit = (np.minimum(q1, q2) <= m) & (m > 0)
            mask |= (hit.astype(np.uint8) << np.uint8(2 * ry + cx))
    return mask


def count(sc, fwd):
    cam = sc.cameras[0]
    W, H = cam.width, cam.height
    gx, gy = (W + 15) // 16, (H + 15) // 16
    co = fwd.conic_opacity
    A2 = (np.float32(-0.5) * LOG2E) * co[:, 0]; B2 = (-LOG2E) * co[:, 1]; C2 = (np.float32(-0.5) * LOG2E) * co[:, 2]
    with np.errstate(divide="ignore"):
        lo = np.log2(co[:, 3]).astype(np.float32)
    mcut = (lo - ALPHA_MIN_LOG2) * np.float32(1.0001) + np.float32(1e-3)
    tot = dict(entries=0, entries_mask0=0, potential=0, warp_visits_nomask=0, warp_visits_mask=0, warp_visits_fwd=0,
               warp_visits_bwd=0, lanes_alpha_fwd=0, lanes_blend_fwd=0, lanes_active_bwd=0, warp_visits_bwd_any=0)
    ncontrib = fwd.n_contrib.reshape(H, W)
    for t in range(gx * gy):
        r0, r1 = int(fwd.ranges[t, 0]), int(fwd.ranges[t, 1])
        n = r1 - r0
        if n == 0:
            continue
        tx, ty = t % gx, t // gx
        ids = fwd.point_list[r0:r1].astype(np.int64)
        mx, my = fwd.mean2D[ids, 0], fwd.mean2D[ids, 1]
        masks = sub_tile_masks(mx, my, -A2[ids], -B2[ids], -C2[ids], mcut[ids], np.float32(16 * tx), np.float32(16 * ty))
        xs = 16 * tx + np.arange(16); ys = 16 * ty + np.arange(16)
        inside = (ys[:, None] < H) & (xs[None, :] < W)                        # [16,16]
        dx = mx[:, None, None] - xs[None, None, :].astype(np.float32)          # [n,1,16]
        dy = my[:, None, None] - ys[None, :, None].astype(np.float32)          # [n,16,1]
        pw = (A2[ids, None, None] * dx) * dx + (B2[ids, None, None] * dy) * dx + (C2[ids, None, None] * dy) * dy
        ee = pw + lo[ids, None, None]
        passes = (pw <= 0) & (ee >= ALPHA_MIN_LOG2) & inside[None]             # alpha >= 1/255
        alpha = np.minimum(np.float32(0.99), np.exp2(np.minimum(ee, 0.0), dtype=np.float32)) * passes
        # transmittance BEFORE each entry, per pixel; the forward stops a pixel at the first entry with T*(1-alpha) < 1e-4
        T_after = np.cumprod(1.0 - alpha.astype(np.float64), axis=0)
        stop_hit = (T_after < 1e-4) & passes
        first_stop = np.where(stop_hit.any(0), stop_hit.argmax(0), n)         # index of the terminating entry, n if none
        alive = np.arange(n)[:, None, None] <= first_stop[None]                # pixel still walking at entry k (incl. the stopper)
        alive &= inside[None]
        blend = passes & (np.arange(n)[:, None, None] < first_stop[None])      # contributing pairs
        # sub-rectangle geometry: warp w = (cx = w&1, ry = w>>1): columns 8cx..8cx+7, rows 4ry..4ry+3
        def per_warp(x):  # [n,16,16] bool -> [n,8] any / [n,8] sum
            v = x.reshape(n, 4, 4, 2, 8)                                       # rows -> (ry, 4), cols -> (cx, 8)
            return v
        wv_alive = per_warp(alive).any(axis=(2, 4))                            # [n,4,2] warp has an undone pixel
        bits = ((masks[:, None, None] >> (2 * np.arange(4)[None, :, None] + np.arange(2)[None, None, :]).astype(np.uint8)) & 1).astype(bool)
        warp_inside = per_warp(np.broadcast_to(inside[None], (n, 16, 16))).any(axis=(2, 4))
        tot["entries"] += n
        tot["entries_mask0"] += int((masks == 0).sum())
        tot["potential"] += int(n * inside.sum())
        tot["warp_visits_nomask"] += int(warp_inside.sum())
        tot["warp_visits_mask"] += int((bits & warp_inside).sum())
        vis_fwd = bits & wv_alive                                              # visits the forward really makes
        tot["warp_visits_fwd"] += int(vis_fwd.sum())
        tot["lanes_alpha_fwd"] += int((per_warp(passes & alive).sum(axis=(2, 4)) * vis_fwd).sum())
        tot["lanes_blend_fwd"] += int(blend.sum())
        # backward: warp walks entries in front of its deepest last contributor; lane active iff k < its own n_contrib and alpha passes
        last = np.zeros((16, 16), np.int64)
        hh, ww = min(16, H - 16 * ty), min(16, W - 16 * tx)
        last[:hh, :ww] = ncontrib[16 * ty:16 * ty + hh, 16 * tx:16 * tx + ww]
        kk = np.arange(n)[:, None, None]
        act = passes & (kk < last[None])
        wmax = last.reshape(4, 4, 2, 8).max(axis=(1, 3))                       # [4,2]
        vis_bwd = bits & (np.arange(n)[:, None, None] < wmax[None])
        tot["warp_visits_bwd"] += int(vis_bwd.sum())
        act_w = per_warp(act).sum(axis=(2, 4))
        tot["lanes_active_bwd"] += int((act_w * vis_bwd).sum())
        tot["warp_visits_bwd_any"] += int(((act_w > 0) & vis_bwd).sum())
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None)
    ap.add_argument("--crop", default="c3:2", help="config:linear_fraction, e.g. c3:2 = 1/4 of c3 at the same density")
    a = ap.parse_args()
    from divshot_b200.scenes import crop_of, make_scene
    from oracle import oracle as orc
    from util import orc_cam, scene_arrays
    if a.config:
        sc, what = make_scene(a.config), a.config
    else:
        name, frac = a.crop.split(":")
        sc, what = crop_of(name, int(frac)), f"1/{int(frac) ** 2} crop of {name}"
    oc = orc_cam(sc.cameras[0], sc.sh_degree)
    fwd = orc.forward(oc, *scene_arrays(sc))
    t = count(sc, fwd)
    D = t["entries"]
    out = {"workload": f"{what}: {sc.N} Gaussians, {sc.cameras[0].width}x{sc.cameras[0].height}", "D": D, "V": int((fwd.radii > 0).sum()),
           "counts": t,
           "ratios": {
               "entries_with_empty_mask": t["entries_mask0"] / D,
               "warp_visits_kept_by_mask": t["warp_visits_mask"] / t["warp_visits_nomask"],
               "fwd_warp_visits_after_early_out": t["warp_visits_fwd"] / t["warp_visits_nomask"],
               "fwd_lanes_passing_alpha_per_visit": t["lanes_alpha_fwd"] / max(1, t["warp_visits_fwd"]),
               "fwd_blended_pairs_over_potential": t["lanes_blend_fwd"] / t["potential"],
               "bwd_warp_visits_over_nomask": t["warp_visits_bwd"] / t["warp_visits_nomask"],
               "bwd_visits_with_an_active_lane": t["warp_visits_bwd_any"] / max(1, t["warp_visits_bwd"]),
               "bwd_active_lanes_per_buffered_visit": t["lanes_active_bwd"] / max(1, t["warp_visits_bwd_any"]),
           }}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
