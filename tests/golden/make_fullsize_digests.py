#!/usr/bin/env python
"""sha256 digests of the oracle's bit-exact integer outputs (radii, tiles_touched, point_list, ranges) at the
BASELINE.json configs c2, c3, c5 -> tests/golden/fullsize_digests.json.  The GPU tests compare the CUDA path's
outputs with these at full size without running the oracle (tests/test_zz_gpu_fullsize.py).  Like every fixture made
from the oracle they do not come from the reference.  Run from the repo root; takes a few minutes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import props  # noqa: E402
from divshot_b200.scenes import make_scene  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from util import orc_cam, scene_arrays  # noqa: E402

if __name__ == "__main__":
    out = {}
    for name in ("c2", "c3", "c5"):
        t0 = time.time()
        sc = make_scene(name, with_grad=False)
        f = orc.forward(orc_cam(sc.cameras[0], sc.sh_degree), *scene_arrays(sc), render=False)
        props.check_binning(f.point_list, f.ranges, f.depth, f.radii, f.mean2D, f.tiles_touched,
                            sc.cameras[0].width, sc.cameras[0].height)
        out[name] = props.index_digests(f.radii, f.tiles_touched, f.point_list, f.ranges)
        print(name, out[name]["D"], out[name]["V"], f"{time.time() - t0:.1f}s", flush=True)
    # the views the multi-GPU runs render (SURVEY.md §8 e): cameras on a ring looking at (0, 0, 6), not the identity view
    for key, name, views, view in (("c3_views8_view1", "c3", 8, 1), ("c3_views8_view5", "c3", 8, 5), ("c4_view3", "c4", 8, 3)):
        t0 = time.time()
        sc = make_scene(name, views=views, with_grad=False)
        cam = sc.cameras[view]
        f = orc.forward(orc_cam(cam, sc.sh_degree), *scene_arrays(sc), render=False)
        props.check_binning(f.point_list, f.ranges, f.depth, f.radii, f.mean2D, f.tiles_touched, cam.width, cam.height)
        out[key] = {**props.index_digests(f.radii, f.tiles_touched, f.point_list, f.ranges), "scene": name, "views": views, "view": view}
        print(key, out[key]["D"], out[key]["V"], f"{time.time() - t0:.1f}s", flush=True)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "fullsize_digests.json"), "w"), indent=1)
