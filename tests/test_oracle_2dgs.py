"""The 2DGS ("surfel", GaussianTrainConfig::modelType = 1) oracle pinned from two independent sides, on the CPU:
closed-form known-answer cases of the published algorithm, and a float64 torch-autograd re-expression of the forward
(tests/autograd_ref_2dgs.py) for the analytic backward.  PARITY UNPINNED by the reference: DIVSHOT's 2DGS rasterizer is in the
closed plugin; the option is documented at docs/userGuide.md:38 and set at main.cpp:28 / gs_train.cpp:68."""
import numpy as np

import autograd_ref_2dgs as ar2
from divshot_b200.scenes import make_scene
from oracle import oracle as orc
from util import assert_close_robust, orc_cam, scene_arrays


def test_single_disk_facing_the_camera_is_a_gaussian_of_the_projected_scale():
    """A disk on the optical axis at depth z, facing the camera, s_u = s_v = s: in pixel space it is the isotropic Gaussian
    exp(-|x - c|^2 / (2 (f s / z)^2)) (f = focal length in pixels), floored by the low-pass filter exp(-|x - c|^2)."""
    W, H = 96, 64
    sc = make_scene(N=1, width=W, height=H, sh_degree=0, seed=1)
    z, s = 4.0, 0.05
    sc.means3D[0] = (0, 0, z); sc.log_scales[0] = np.log([s, s, 1e-3]); sc.quats[0] = (1, 0, 0, 0); sc.logit_opac[:] = 3.0
    sc.sh0[0] = (1.0, 0.5, -0.2)
    cam = sc.cameras[0]
    f = orc.forward2d(orc_cam(cam, 0), *scene_arrays(sc))
    fpx = W / (2 * cam.tanfovx)
    sig = fpx * s / z
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    assert np.allclose(f.mean2D[0], (cx, cy), atol=1e-3)
    assert f.radii[0] == int(np.ceil(max(3 * sig, 3 * 0.707106)))
    ys, xs = np.mgrid[0:H, 0:W]
    r2 = (xs - cx) ** 2 + (ys - cy) ** 2
    o = 1 / (1 + np.exp(-3.0))
    alpha = np.minimum(0.99, o * np.exp(-0.5 * np.minimum(r2 / sig ** 2, 2 * r2)))
    alpha[alpha < 1 / 255] = 0
    col = np.maximum(0.28209479177387814 * sc.sh0[0] + 0.5, 0)
    want = alpha[None] * col[:, None, None]
    inside = r2 <= (f.radii[0] - 1) ** 2   # (outside the tile rect of the radius nothing is drawn at all)
    assert np.abs(f.image - want)[:, inside].max() < 2e-5
    assert np.allclose(f.final_T.reshape(H, W)[inside], (1 - alpha)[inside], atol=2e-6)


def test_tilted_disk_foreshortens_and_the_third_scale_is_ignored():
    W, H = 128, 96
    sc = make_scene(N=1, width=W, height=H, sh_degree=0, seed=2)
    sc.means3D[0] = (0.1, -0.05, 5.0); sc.log_scales[0] = np.log([0.2, 0.2, 0.5]); sc.logit_opac[:] = 2.0
    th = np.deg2rad(60)                      # rotate about the y axis: the disk's u axis tilts away from the image plane
    sc.quats[0] = (np.cos(th / 2), 0, np.sin(th / 2), 0)
    f1 = orc.forward2d(orc_cam(sc.cameras[0], 0), *scene_arrays(sc))
    sc.log_scales[0, 2] = np.log(5.0)
    f2 = orc.forward2d(orc_cam(sc.cameras[0], 0), *scene_arrays(sc))
    assert np.array_equal(f1.image, f2.image) and np.array_equal(f1.radii, f2.radii)
    a = 1 - f1.final_T.reshape(H, W)
    cy, cx = np.unravel_index(np.argmax(a), a.shape)
    wx = (a[cy] > 0.5 * a.max()).sum(); wy = (a[:, cx] > 0.5 * a.max()).sum()
    assert 0.4 < wx / wy < 0.62, (wx, wy)     # cos(60 deg) = 0.5: half as wide as tall


def test_analytic_backward_matches_float64_autograd():
    sc = make_scene(N=1200, width=64, height=48, sh_degree=2, seed=5, normalise_quats=False, bg=(0.2, 0.4, 0.1))
    sc.log_scales += 1.0
    oc = orc_cam(sc.cameras[0], 2)
    f = orc.forward2d(oc, *scene_arrays(sc))
    b = orc.backward2d(oc, f, *scene_arrays(sc), sc.dL_dpix[0])
    assert f.D > 2000 and (f.radii > 0).mean() > 0.5
    img, g, proj, n_contrib, final_T = ar2.render_and_grad_2d(sc.cameras[0], scene_arrays(sc), 2, f.ranges, f.point_list, f.radii,
                                                             sc.dL_dpix[0])
    ok = (f.fragile == 0).reshape(48, 64)
    assert ok.mean() > 0.9
    assert np.abs(f.image - img)[:, ok].max() < 1e-4
    assert (f.n_contrib.reshape(48, 64)[ok] == n_contrib[ok]).all()
    for name, ref in [("dL_dmeans3D", g["means3D"]), ("dL_dscales", g["scales"]), ("dL_dquats", g["quats"]),
                      ("dL_dopacities", g["opac"]), ("dL_dsh0", g["sh0"]), ("dL_dshN", g["shN"])]:
        got = getattr(b, name)
        assert_close_robust(got, ref.reshape(got.shape), 1e-4, name)
    assert not b.dL_dscales[:, 2].any(), "the third scale of a surfel has no gradient"
