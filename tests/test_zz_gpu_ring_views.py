"""The views the multi-GPU runs render — cameras on a ring looking at (0, 0, 6), i.e. rotated
and translated views, not the identity camera of c3 / c5 — at FULL size (1 M Gaussians, 1600x1000): the bit-exact index
outputs against digests of the oracle (tests/golden/fullsize_digests.json, made by make_fullsize_digests.py) and the
size-independent properties of tests/props.py.  SURVEY.md §8 e: "tile/sort indices are per-view and stay bit-exact"."""
import json
import os

import pytest

import props
from divshot_b200.scenes import make_scene

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("key", ["c3_views8_view1", "c3_views8_view5", "c4_view3"])
def test_ring_view_indices_at_full_size(key):
    from divshot_b200 import _cabi
    from divshot_b200.rasterizer import Rasterizer, scene_to_device
    gold = json.load(open(os.path.join(HERE, "golden", "fullsize_digests.json")))[key]
    sc = make_scene(gold["scene"], views=gold["views"], with_grad=False)
    cam = sc.cameras[gold["view"]]
    r = Rasterizer(0)
    try:
        params = scene_to_device(sc, r.device)
        dcam = _cabi.make_camera(cam, sc.sh_degree)
        img, radii = r.forward(dcam, params)
        rd = r.debug_read
        tt, pl, rg = rd(_cabi.BUF_TILES_TOUCHED), rd(_cabi.BUF_POINT_LIST), rd(_cabi.BUF_RANGES)
        got = props.index_digests(radii.cpu().numpy(), tt, pl, rg)
        assert (got["D"], got["V"]) == (gold["D"], gold["V"]), (got["D"], got["V"], gold["D"], gold["V"])
        for k in ("radii", "tiles_touched", "ranges", "point_list"):
            assert got[k] == gold[k], f"{key}: {k} differs from the oracle at full size"
        props.check_binning(pl, rg, rd(_cabi.BUF_DEPTH), radii.cpu().numpy(), rd(_cabi.BUF_MEAN2D), tt, cam.width, cam.height)
        props.check_compositing(img.cpu().numpy(), rd(_cabi.BUF_FINAL_T), rd(_cabi.BUF_N_CONTRIB), rg, cam.width, cam.height)
    finally:
        r.close()
