"""BASELINE.json's full-size configurations on the GPU (c3: 1 M Gaussians @ 1600x1000, the bench workload; c5: 5 M @
1920x1080), where comparing every float with a live oracle run is too slow for a test:

  * the bit-exact integer outputs (radii, tiles_touched, point_list, ranges) against sha256 digests of the ORACLE's
    outputs frozen in tests/golden/fullsize_digests.json (tests/golden/make_fullsize_digests.py);
  * the size-independent properties of tests/props.py (validated on oracle outputs and on corrupted copies by the CPU
    test tests/test_props.py): ranges partition the list in tile order, every tile list strictly sorted by
    (depth bits, id), per-tile lengths = histogram of the tile rectangles rebuilt from (mean2D, radius), every listed
    Gaussian covers its tile, final_T in [1e-4, 1], n_contrib within the list, background-only pixels exact;
  * idempotence of the forward, equality of the two binning modes (two-pass / single-pass deferred-check) at full size;
  * the backward: finite, exactly zero for invisible Gaussians, linear in dL/dpixel (c3).
The checks themselves live in tests/fullsize_checks.py; the CPU suite runs the same code on an oracle-backed stand-in.
"""
import json
import os

import pytest

import fullsize_checks
from divshot_b200.scenes import make_scene

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", ["c3", "c5"])
def test_fullsize_config(name):
    from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
    gold = json.load(open(os.path.join(HERE, "golden", "fullsize_digests.json")))[name]
    sc = make_scene(name)
    r = Rasterizer(0)
    try:
        fullsize_checks.run(r, sc, scene_to_device(sc, r.device), gold, linearity=(name == "c3"),
                            grad_alloc=GradBuffers.allocate)
    finally:
        r.close()
