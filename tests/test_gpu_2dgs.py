"""GPU parity of the 2DGS ("surfel") variant — GaussianTrainConfig::modelType = 1 (main.cpp:28, gs_train.cpp:68,
docs/userGuide.md:38), DVS_FLAG_MODEL_2DGS of the C-ABI — against the CPU oracle (oracle/dvs_oracle.c, S.1-S.4; itself pinned
by tests/test_oracle_2dgs.py).  Same bar as the 3DGS path: index outputs bit-exact, image and gradients within 1e-4.
Parity unpinned by the reference (its 2DGS rasterizer is in the closed plugin)."""
import os
import subprocess

import numpy as np
import pytest
import torch

from divshot_b200 import _cabi
from divshot_b200.scenes import make_scene
from oracle import oracle as orc
from util import assert_close, assert_close_robust, check_image_against_oracle, orc_cam, scene_arrays

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rast():
    from divshot_b200.rasterizer import Rasterizer
    r = Rasterizer(0)
    yield r
    r.close()


def _run2d(rast, sc, deg, defer=False):
    from divshot_b200.rasterizer import GradBuffers, scene_to_device
    dev = rast.device
    params = scene_to_device(sc, dev)
    cam = _cabi.make_camera(sc.cameras[0], deg, sh_rest_alloc=sc.shN.shape[1], flags=_cabi.FLAG_MODEL_2DGS)
    if defer:
        rast.forward(cam, params); rast.forward(cam, params)
    img, radii = rast.forward(cam, params, defer_check=defer)
    out = dict(image=img.cpu().numpy(), radii=radii.cpu().numpy(), stats=rast.stats())
    for name, which in [("tiles_touched", _cabi.BUF_TILES_TOUCHED), ("depth", _cabi.BUF_DEPTH), ("mean2D", _cabi.BUF_MEAN2D),
                        ("rgb", _cabi.BUF_RGB), ("clamped", _cabi.BUF_CLAMPED), ("point_list", _cabi.BUF_POINT_LIST),
                        ("ranges", _cabi.BUF_RANGES), ("final_T", _cabi.BUF_FINAL_T), ("n_contrib", _cabi.BUF_N_CONTRIB)]:
        out[name] = rast.debug_read(which)
    g = GradBuffers.allocate(sc.N, sc.shN.shape[1], dev)
    g.flat.fill_(float("nan"))
    m2 = torch.zeros(sc.N, 2, device=dev)
    rast.backward(torch.from_numpy(sc.dL_dpix[0]).to(dev), g, mean2D=m2)
    torch.cuda.synchronize()
    out["grads"] = {k: getattr(g, k).cpu().numpy() for k in ("means3D", "scales", "quats", "opacities", "sh0", "shN")}
    out["mean2D_grad"] = m2.cpu().numpy()
    return out


def _check(sc, got, deg):
    oc = orc_cam(sc.cameras[0], deg, sh_rest_alloc=sc.shN.shape[1])
    f = orc.forward2d(oc, *scene_arrays(sc))
    b = orc.backward2d(oc, f, *scene_arrays(sc), sc.dL_dpix[0])
    vis = f.radii > 0
    assert vis.mean() > 0.3 and f.D > sc.N // 2
    assert np.array_equal(got["radii"], f.radii), "radii"
    assert np.array_equal(got["tiles_touched"], f.tiles_touched), "tiles_touched"
    assert np.array_equal(got["depth"].view(np.uint32)[vis], f.depth.view(np.uint32)[vis]), "depth bits"
    assert np.array_equal(got["mean2D"].view(np.uint32)[vis], f.mean2D.view(np.uint32)[vis]), "projected centre bits"
    assert np.array_equal(got["rgb"].view(np.uint32)[vis], f.rgb.view(np.uint32)[vis]), "rgb bits"
    assert np.array_equal(got["clamped"][vis], f.clamped[vis])
    assert got["stats"]["num_dups"] == f.D and got["stats"]["num_visible"] == int(vis.sum())
    assert np.array_equal(got["ranges"], f.ranges) and np.array_equal(got["point_list"], f.point_list), "sorted tile lists"
    check_image_against_oracle(got["image"], got["final_T"], got["n_contrib"], f, min_robust=0.85)
    for k, ref in [("means3D", b.dL_dmeans3D), ("scales", b.dL_dscales), ("quats", b.dL_dquats), ("opacities", b.dL_dopacities),
                   ("sh0", b.dL_dsh0), ("shN", b.dL_dshN)]:
        a = got["grads"][k]
        assert np.isfinite(a).all(), f"{k}: non-finite / unwritten gradient"
        if ref.size:
            assert_close_robust(a, ref.reshape(a.shape), 1e-4, f"2DGS dL_d{k}")
    assert not got["grads"]["scales"][:, 2].any(), "the third scale of a surfel has no gradient"
    W, H = sc.cameras[0].width, sc.cameras[0].height
    assert_close_robust(got["mean2D_grad"], b.dL_dmean2D * np.array([0.5 * W, 0.5 * H], np.float32), 1e-4, "dL_dmean2D (ndc-scaled)")
    return f


@pytest.mark.parametrize("deg,N,W,H,seed", [(0, 3000, 96, 64, 211), (1, 5000, 131, 77, 212), (3, 6000, 160, 96, 214)])
def test_2dgs_small_scenes(rast, deg, N, W, H, seed):
    sc = make_scene(N=N, width=W, height=H, sh_degree=deg, seed=seed, normalise_quats=False, bg=(0.2, 0.5, 0.1))
    sc.log_scales += 1.0
    _check(sc, _run2d(rast, sc, deg), deg)


def test_2dgs_deferred_check_mode_and_3dgs_afterwards(rast):
    """The training-loop mode (no host synchronisation) gives the same 2DGS result; a 3DGS forward on the same context
    afterwards is a 3DGS forward again (the backward follows the forward's model)."""
    from divshot_b200.rasterizer import GradBuffers, scene_to_device
    sc = make_scene(N=4000, width=112, height=80, sh_degree=1, seed=221, normalise_quats=False)
    sc.log_scales += 1.0
    _check(sc, _run2d(rast, sc, 1, defer=True), 1)
    params = scene_to_device(sc, rast.device)
    cam = _cabi.make_camera(sc.cameras[0], 1)
    img, _ = rast.forward(cam, params)
    g = GradBuffers.allocate(sc.N, 3, rast.device)
    rast.backward(torch.from_numpy(sc.dL_dpix[0]).to(rast.device), g)
    oc = orc_cam(sc.cameras[0], 1)
    f = orc.forward(oc, *scene_arrays(sc))
    b = orc.backward(oc, f, *scene_arrays(sc), sc.dL_dpix[0])
    check_image_against_oracle(img.cpu().numpy(), rast.debug_read(_cabi.BUF_FINAL_T), rast.debug_read(_cabi.BUF_N_CONTRIB), f)
    assert_close_robust(g.means3D.cpu().numpy(), b.dL_dmeans3D, 1e-4, "3DGS after 2DGS: dL_dmeans3D")


def test_2dgs_trains_through_the_plugin(tmp_path):
    """--modelType 1 through the plugin boundary (the caller's loop of gs_train.cpp:152-167): targets rendered as surfels,
    the loss falls by more than 20 % (the driver's exit code)."""
    from divshot_b200 import build
    libs = build.build_all(torch_binding=False)
    lib_dir = os.path.join(ROOT, "divshot_b200", "lib")
    r = subprocess.run([libs["gstrain_driver"], "synthetic:N=20000,W=320,H=240,views=4,deg=1", "300", str(tmp_path / "m2d.ply"), "modelType=1"],
                       capture_output=True, text=True, env={**os.environ, "LD_LIBRARY_PATH": lib_dir}, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


def test_2dgs_cull_masks_are_conservative(rast):
    """Every (pixel, surfel) pair the oracle's rule blends (alpha >= 1/255 after the S.2 tests) lies in an 8x4 sub-rectangle
    whose mask bit is set — the compositor skips whole warps on that bit, so a missing bit would silently drop light."""
    from divshot_b200.rasterizer import scene_to_device
    sc = make_scene(N=2500, width=96, height=64, sh_degree=0, seed=231, normalise_quats=False)
    sc.log_scales += 1.2
    sc.logit_opac[::3] += 3.0   # some nearly opaque surfels: widest alpha support
    sc.logit_opac[1::7] -= 6.0  # some below 1/255: empty masks
    params = scene_to_device(sc, rast.device)
    cam = _cabi.make_camera(sc.cameras[0], 0, flags=_cabi.FLAG_MODEL_2DGS)
    rast.forward(cam, params)
    mask = rast.debug_read(_cabi.BUF_CULL_MASK)
    f = orc.forward2d(orc_cam(sc.cameras[0], 0), *scene_arrays(sc), render=False)
    assert np.array_equal(rast.debug_read(_cabi.BUF_POINT_LIST), f.point_list)
    W, H = 96, 64
    gx = (W + 15) // 16
    bad = kept = total = 0
    for tile in range(f.ranges.shape[0]):
        r0, r1 = (int(v) for v in f.ranges[tile])
        x0, y0 = (tile % gx) * 16, (tile // gx) * 16
        ys, xs = np.mgrid[y0:y0 + 16, x0:x0 + 16].astype(np.float64)
        for j in range(r0, r1):
            g = int(f.point_list[j])
            T = f.transmat[g].astype(np.float64)
            Tu, Tv, Tw = T[0:3], T[3:6], T[6:9]
            k = xs[..., None] * Tw - Tu; l = ys[..., None] * Tw - Tv
            pv = np.cross(k, l)
            with np.errstate(divide="ignore", invalid="ignore"):
                u, v = pv[..., 0] / pv[..., 2], pv[..., 1] / pv[..., 2]
            rho3d = u * u + v * v
            rho2d = 2.0 * ((f.mean2D[g, 0] - xs) ** 2 + (f.mean2D[g, 1] - ys) ** 2)
            use3d = rho3d <= rho2d
            dep = np.where(use3d, u * Tw[0] + v * Tw[1] + Tw[2], Tw[2])
            alpha = np.minimum(0.99, f.opacity[g] * np.exp(-0.5 * np.minimum(rho3d, rho2d)))
            contrib = (pv[..., 2] != 0) & (dep >= 0.2) & (alpha >= 1 / 255) & (xs < W) & (ys < H)
            sub = contrib.reshape(4, 4, 2, 8).any(axis=(1, 3))  # [row(4), col(2)]
            m = int(mask[j])
            total += 8; kept += bin(m).count("1")
            for r in range(4):
                for c in range(2):
                    if sub[r, c] and not (m >> (2 * r + c)) & 1:
                        bad += 1
    assert bad == 0, f"{bad} contributing sub-rectangles without their mask bit"
    assert kept < 0.8 * total, f"the masks cull nothing ({kept} of {total} bits set)"
