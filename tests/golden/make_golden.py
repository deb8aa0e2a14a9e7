#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no implementation, test or golden vector for this path, so these
fixtures do NOT come from the reference: they freeze the oracle (which restates the credited public algorithm and
is itself pinned by closed-form KATs and a float64 autograd re-expression).  They make any later change to the
oracle's results visible, and let the GPU parity tests compare against files instead of a live oracle run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from divshot_b200.scenes import make_scene  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from util import orc_cam, scene_arrays  # noqa: E402

CASES = {
    # name: (make_scene kwargs, log-scale shift)
    "g8_small_deg3": (dict(N=3000, width=96, height=64, sh_degree=3, seed=11, normalise_quats=False, bg=(0.2, 0.5, 0.1)), 0.8),
    "g8_c1": (dict(name="c1"), 0.0),
}


def build(name):
    kw, shift = CASES[name]
    kw = dict(kw)
    sc = make_scene(kw.pop("name", None), **kw)
    sc.log_scales += shift
    return sc


def run(name):
    sc = build(name)
    oc = orc_cam(sc.cameras[0], sc.sh_degree)
    f = orc.forward(oc, *scene_arrays(sc), threads=1)
    b = orc.backward(oc, f, *scene_arrays(sc), sc.dL_dpix[0], threads=1)
    return dict(radii=f.radii, tiles_touched=f.tiles_touched, depth_bits=f.depth.view(np.uint32), point_list=f.point_list,
                ranges=f.ranges, n_contrib=f.n_contrib, fragile=f.fragile, image=f.image, final_T=f.final_T,
                dL_dmeans3D=b.dL_dmeans3D, dL_dscales=b.dL_dscales, dL_dquats=b.dL_dquats,
                dL_dopacities=b.dL_dopacities, dL_dsh0=b.dL_dsh0, dL_dshN=b.dL_dshN, dL_dmean2D=b.dL_dmean2D)


if __name__ == "__main__":
    for name in CASES:
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
        np.savez_compressed(out, **run(name))
        print(name, os.path.getsize(out) // 1024, "KiB")
