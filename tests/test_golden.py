"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle — the reference holds none)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402
from util import assert_close_robust, elem_err  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_reproduces_golden(name):
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    now = mg.run(name)
    for k in ("radii", "tiles_touched", "depth_bits", "point_list", "ranges", "n_contrib", "fragile"):
        assert np.array_equal(now[k], gold[k]), k
    for k in ("image", "final_T", "dL_dmeans3D", "dL_dscales", "dL_dquats", "dL_dopacities", "dL_dsh0", "dL_dshN", "dL_dmean2D"):
        assert np.array_equal(now[k], gold[k]), k  # single-threaded oracle is deterministic


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_cuda_matches_golden(name):
    import torch
    from divshot_b200 import _cabi
    from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    sc = mg.build(name)
    r = Rasterizer(0)
    try:
        params = scene_to_device(sc, r.device)
        cam = _cabi.make_camera(sc.cameras[0], sc.sh_degree)
        img, radii = r.forward(cam, params)
        g = GradBuffers.allocate(sc.N, sc.shN.shape[1], r.device)
        m2 = torch.zeros(sc.N, 2, device=r.device)
        r.backward(torch.from_numpy(sc.dL_dpix[0]).to(r.device), g, mean2D=m2)
        assert np.array_equal(radii.cpu().numpy(), gold["radii"])
        assert np.array_equal(r.debug_read(_cabi.BUF_TILES_TOUCHED), gold["tiles_touched"])
        assert np.array_equal(r.debug_read(_cabi.BUF_POINT_LIST), gold["point_list"])
        assert np.array_equal(r.debug_read(_cabi.BUF_RANGES), gold["ranges"])
        vis = gold["radii"] > 0
        assert np.array_equal(r.debug_read(_cabi.BUF_DEPTH).view(np.uint32)[vis], gold["depth_bits"][vis])
        ok = gold["fragile"] == 0
        assert np.array_equal(r.debug_read(_cabi.BUF_N_CONTRIB)[ok], gold["n_contrib"][ok])
        e = elem_err(img.cpu().numpy(), gold["image"]).reshape(3, -1)
        assert e[:, ok].max() <= 1e-4 and e.max() <= 1e-2
        for k, t in [("dL_dmeans3D", g.means3D), ("dL_dscales", g.scales), ("dL_dquats", g.quats),
                     ("dL_dopacities", g.opacities), ("dL_dsh0", g.sh0), ("dL_dshN", g.shN), ("dL_dmean2D", m2)]:
            if gold[k].size:
                assert_close_robust(t.cpu().numpy(), gold[k].reshape(t.shape), 1e-4, k)
    finally:
        r.close()
