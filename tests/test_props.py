"""The size-independent property checkers of tests/props.py, exercised on the oracle's outputs (they must accept
them) and on corrupted copies (they must reject them) — so that the GPU tests can rely on them at BASELINE.json's
full sizes (tests/test_zz_gpu_fullsize.py).  Also freezes/validates the index digests of the full-size configs."""
import json
import os

import numpy as np
import pytest

import props
from divshot_b200.scenes import make_scene
from oracle import oracle as orc
from util import orc_cam, scene_arrays

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def case():
    sc = make_scene(N=20000, width=200, height=120, sh_degree=1, seed=77, normalise_quats=False, bg=(0.1, 0.0, 0.3))
    sc.log_scales += 0.9
    oc = orc_cam(sc.cameras[0], 1)
    f = orc.forward(oc, *scene_arrays(sc))
    return sc, oc, f


def _bin_args(sc, f):
    return dict(point_list=f.point_list, ranges=f.ranges, depth=f.depth, radii=f.radii, mean2D=f.mean2D,
                tiles_touched=f.tiles_touched, width=sc.cameras[0].width, height=sc.cameras[0].height)


def test_oracle_outputs_satisfy_every_property(case):
    sc, oc, f = case
    assert f.D > 50000 and (f.radii == 0).sum() > 100
    rect, _, _ = props.tile_rects(f.mean2D, f.radii, 200, 120)
    assert np.array_equal(rect, f.rect), "rectangles recomputed from (mean2D, radius) must equal the oracle's"
    props.check_binning(**_bin_args(sc, f))
    props.check_compositing(f.image, f.final_T, f.n_contrib, f.ranges, 200, 120, bg=(0.1, 0.0, 0.3))
    rng = np.random.default_rng(1)
    u = rng.normal(size=(3, 120, 200)).astype(np.float32)
    v = rng.normal(size=(3, 120, 200)).astype(np.float32)
    names = ("dL_dmeans3D", "dL_dscales", "dL_dquats", "dL_dopacities", "dL_dsh0", "dL_dshN")
    g = [{k: getattr(orc.backward(oc, f, *scene_arrays(sc), d), k) for k in names} for d in (u, v, 2.0 * u - 0.5 * v)]
    props.check_gradient_support(g[0], f.radii)
    props.check_backward_linearity(g[0], g[1], g[2], 2.0, -0.5)


def test_checkers_reject_corrupted_outputs(case):
    sc, oc, f = case
    a = _bin_args(sc, f)
    t = int(np.argmax(f.ranges[:, 1] - f.ranges[:, 0] > 3))
    s = int(f.ranges[t, 0])

    def corrupt(**kw):
        b = dict(a)
        b.update(kw)
        with pytest.raises(AssertionError):
            props.check_binning(**b)

    pl = f.point_list.copy(); pl[s], pl[s + 1] = pl[s + 1], pl[s]
    corrupt(point_list=pl)                                             # two neighbours swapped: order broken
    pl = f.point_list.copy(); pl[s + 1] = pl[s]
    corrupt(point_list=pl)                                             # duplicate id inside one tile
    rg = f.ranges.copy(); rg[t, 1] -= 1
    corrupt(ranges=rg)                                                 # a tile lost an entry
    tt = f.tiles_touched.copy(); tt[np.argmax(f.radii > 0)] += 1
    corrupt(tiles_touched=tt)                                          # count does not match the rectangle
    far = int(np.argmax((f.radii > 0) & (np.abs(f.mean2D[:, 0] - f.mean2D[f.point_list[s], 0]) > 100)))
    pl = f.point_list.copy(); pl[s] = far
    corrupt(point_list=pl)                                             # an id whose rectangle misses the tile
    with pytest.raises(AssertionError):
        ft = f.final_T.copy(); ft[5] = 5e-5
        props.check_compositing(f.image, ft, f.n_contrib, f.ranges, 200, 120, bg=(0.1, 0.0, 0.3))
    with pytest.raises(AssertionError):
        nc = f.n_contrib.copy(); nc[7] = 10 ** 6
        props.check_compositing(f.image, f.final_T, nc, f.ranges, 200, 120, bg=(0.1, 0.0, 0.3))
    with pytest.raises(AssertionError):
        props.check_gradient_support({"x": np.ones((sc.N, 3), np.float32)}, f.radii)
    with pytest.raises(AssertionError):
        g = {"x": np.ones(4)}
        props.check_backward_linearity(g, g, {"x": 2.2 * np.ones(4)}, 1.0, 1.0)


def test_empty_scene_and_single_gaussian():
    for N in (0, 1):
        sc = make_scene(N=max(N, 1), width=40, height=24, sh_degree=0, seed=3)
        arrays = [a[:N] for a in scene_arrays(sc)]
        oc = orc_cam(sc.cameras[0], 0)
        f = orc.forward(oc, *arrays)
        props.check_binning(f.point_list, f.ranges, f.depth, f.radii, f.mean2D, f.tiles_touched, 40, 24)
        props.check_compositing(f.image, f.final_T, f.n_contrib, f.ranges, 40, 24)


@pytest.mark.parametrize("name", ["c2"])
def test_fullsize_index_digests_are_reproduced_by_the_oracle(name):
    """tests/golden/fullsize_digests.json (made by tests/golden/make_fullsize_digests.py) freezes the oracle's integer
    outputs at the BASELINE configs; c2 is re-derived here on every CPU run, c3/c5 only by the generator (minutes)."""
    gold = json.load(open(os.path.join(HERE, "golden", "fullsize_digests.json")))
    sc = make_scene(name)
    f = orc.forward(orc_cam(sc.cameras[0], sc.sh_degree), *scene_arrays(sc), render=False)
    assert props.index_digests(f.radii, f.tiles_touched, f.point_list, f.ranges) == gold[name]


class _OracleRasterizer:
    """Stand-in with the Rasterizer surface tests/fullsize_checks.py uses, backed by the oracle (CPU tensors)."""

    def __init__(self, sc):
        import torch
        self.sc, self.device, self.torch = sc, torch.device("cpu"), torch
        self.f = None

    def forward(self, cam, params, defer_check=False):
        assert cam.width == self.sc.cameras[0].width and cam.sh_degree == self.sc.sh_degree
        self.oc = orc_cam(self.sc.cameras[0], self.sc.sh_degree)
        self.f = orc.forward(self.oc, *scene_arrays(self.sc))
        return self.torch.from_numpy(self.f.image.copy()), self.torch.from_numpy(self.f.radii.copy())

    def backward(self, dl, g):
        b = orc.backward(self.oc, self.f, *scene_arrays(self.sc), dl.numpy())
        for k, v in (("means3D", b.dL_dmeans3D), ("scales", b.dL_dscales), ("quats", b.dL_dquats),
                     ("opacities", b.dL_dopacities), ("sh0", b.dL_dsh0), ("shN", b.dL_dshN)):
            getattr(g, k).copy_(self.torch.from_numpy(np.ascontiguousarray(v)).reshape(getattr(g, k).shape))

    def debug_read(self, which):
        from divshot_b200 import _cabi
        f = self.f
        return {_cabi.BUF_TILES_TOUCHED: f.tiles_touched, _cabi.BUF_POINT_LIST: f.point_list, _cabi.BUF_RANGES: f.ranges,
                _cabi.BUF_DEPTH: f.depth, _cabi.BUF_MEAN2D: f.mean2D, _cabi.BUF_FINAL_T: f.final_T,
                _cabi.BUF_N_CONTRIB: f.n_contrib}[which]

    def stats(self):
        return {"overflow": 0}


def test_fullsize_check_body_runs_on_an_oracle_stand_in():
    """tests/fullsize_checks.run is what the GPU executes at c3 / c5 (tests/test_zz_gpu_fullsize.py); here the same code
    runs at c2 with the oracle standing in for the device, so a mistake in the checks shows up without a GPU."""
    import fullsize_checks
    from divshot_b200.rasterizer import GradBuffers
    gold = json.load(open(os.path.join(HERE, "golden", "fullsize_digests.json")))["c2"]
    sc = make_scene("c2")
    fullsize_checks.run(_OracleRasterizer(sc), sc, None, gold, linearity=True, grad_alloc=GradBuffers.allocate)
    bad = dict(gold, point_list="0" * 64)
    with pytest.raises(AssertionError, match="point_list differs"):
        fullsize_checks.run(_OracleRasterizer(sc), sc, None, bad, linearity=False, grad_alloc=GradBuffers.allocate)
