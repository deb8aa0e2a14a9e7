"""G7: the oracle's analytic backward vs an independent float64 autograd re-expression."""
import numpy as np
import pytest

import autograd_ref as ar
from divshot_b200.scenes import make_scene
from oracle import oracle as orc
from util import assert_close, assert_close_robust, orc_cam, scene_arrays


@pytest.mark.parametrize("deg,seed,bg,normq", [(0, 1, (0, 0, 0), True), (1, 2, (0.3, 0.1, 0.7), False),
                                                 (3, 3, (0, 0, 0), False), (2, 4, (1, 1, 1), True)])
def test_oracle_backward_matches_autograd(deg, seed, bg, normq):
    sc = make_scene(N=1500, width=64, height=48, sh_degree=deg, seed=seed, normalise_quats=normq, bg=bg)
    # make splats big enough to overlap many pixels and saturate some of them
    sc.log_scales += 1.2
    if seed == 2:  # push some splats beyond 1.3*tanfov: exercises the EWA clamp path (G4)
        sc.means3D[:, :2] *= 1.3
    cam = sc.cameras[0]
    oc = orc_cam(cam, deg)
    fwd = orc.forward(oc, *scene_arrays(sc), threads=1)
    bwd = orc.backward(oc, fwd, *scene_arrays(sc), sc.dL_dpix[0], threads=1)
    img, g, proj, n_contrib, final_T = ar.render_and_grad(cam, scene_arrays(sc), deg, fwd.ranges, fwd.point_list,
                                                          fwd.radii, sc.dL_dpix[0])
    assert fwd.D > 0 and (fwd.n_contrib > 0).mean() > 0.5
    # forward agrees (fp32 vs fp64)
    vis = fwd.radii > 0
    assert_close(fwd.mean2D[vis], proj["mean2D"].detach().numpy()[vis], 2e-5, "mean2D")
    assert_close(fwd.conic_opacity[vis, :3], proj["conic"].detach().numpy()[vis], 1e-4, "conic")
    assert_close(fwd.rgb[vis], proj["rgb"].detach().numpy()[vis], 1e-5, "rgb")
    assert_close(fwd.image, img, 1e-4, "image")
    ok = fwd.fragile.reshape(cam.height, cam.width) == 0
    assert (fwd.n_contrib.reshape(cam.height, cam.width)[ok] == n_contrib[ok]).all()
    # backward agrees
    # small, faint splats: a few near-threshold alpha decisions differ between fp32 and fp64 -> robust metric
    assert_close_robust(bwd.dL_dmeans3D, g["means3D"], 1e-4, "dL_dmeans3D", frac=0.995)
    assert_close_robust(bwd.dL_dscales, g["scales"], 1e-4, "dL_dscales", frac=0.995)
    assert_close_robust(bwd.dL_dquats, g["quats"], 1e-4, "dL_dquats", frac=0.995)
    assert_close_robust(bwd.dL_dopacities, g["opac"].reshape(-1), 1e-4, "dL_dopacity", frac=0.995)
    assert_close(bwd.dL_dsh0, g["sh0"], 2e-4, "dL_dsh0")
    if deg > 0:
        assert_close(bwd.dL_dshN, g["shN"], 2e-4, "dL_dshN")


def test_oracle_backward_activated_inputs():
    sc = make_scene(N=800, width=48, height=48, sh_degree=1, seed=9)
    sc.log_scales += 1.0
    cam = sc.cameras[0]
    act = (sc.means3D, np.exp(sc.log_scales), sc.quats, 1 / (1 + np.exp(-sc.logit_opac)), sc.sh0, sc.shN)
    oc = orc_cam(cam, 1, flags=orc.FLAG_INPUT_ACTIVATED)
    fwd = orc.forward(oc, *act, threads=1)
    bwd = orc.backward(oc, fwd, *act, sc.dL_dpix[0], threads=1)
    img, g, *_ = ar.render_and_grad(cam, act, 1, fwd.ranges, fwd.point_list, fwd.radii, sc.dL_dpix[0],
                                    activated=True)
    assert_close(fwd.image, img, 1e-4, "image")
    # small, faint splats: a few near-threshold alpha decisions differ between fp32 and fp64 -> robust metric
    assert_close_robust(bwd.dL_dmeans3D, g["means3D"], 1e-4, "dL_dmeans3D", frac=0.995)
    assert_close_robust(bwd.dL_dscales, g["scales"], 1e-4, "dL_dscales", frac=0.995)
    assert_close_robust(bwd.dL_dquats, g["quats"], 1e-4, "dL_dquats", frac=0.995)
    assert_close_robust(bwd.dL_dopacities, g["opac"].reshape(-1), 1e-4, "dL_dopacity", frac=0.995)


def test_oracle_backward_antialias_flag():
    """mipAntiliased: opacity compensation sqrt(max(0, det(S')/det(S'+0.3I))) and its gradient into cov2D."""
    sc = make_scene(N=1200, width=64, height=48, sh_degree=1, seed=17, normalise_quats=False)
    sc.log_scales += 0.3  # small splats: the compensation factor is well below 1
    cam = sc.cameras[0]
    oc = orc_cam(cam, 1, flags=orc.FLAG_ANTIALIAS)
    fwd = orc.forward(oc, *scene_arrays(sc), threads=1)
    bwd = orc.backward(oc, fwd, *scene_arrays(sc), sc.dL_dpix[0], threads=1)
    img, g, proj, *_ = ar.render_and_grad(cam, scene_arrays(sc), 1, fwd.ranges, fwd.point_list, fwd.radii, sc.dL_dpix[0],
                                          antialias=True)
    vis = fwd.radii > 0
    plain = orc.forward(orc_cam(cam, 1), *scene_arrays(sc), threads=1, render=False)
    assert (fwd.conic_opacity[vis, 3] < 0.98 * plain.conic_opacity[vis, 3]).mean() > 0.3  # the flag does something
    assert_close(fwd.conic_opacity[vis, 3], proj["opacity"].detach().numpy()[vis], 1e-4, "compensated opacity")
    assert_close(fwd.image, img, 1e-4, "image")
    # small, faint splats: a few near-threshold alpha decisions differ between fp32 and fp64 -> robust metric
    assert_close_robust(bwd.dL_dmeans3D, g["means3D"], 1e-4, "dL_dmeans3D", frac=0.995)
    assert_close_robust(bwd.dL_dscales, g["scales"], 1e-4, "dL_dscales", frac=0.995)
    assert_close_robust(bwd.dL_dquats, g["quats"], 1e-4, "dL_dquats", frac=0.995)
    assert_close_robust(bwd.dL_dopacities, g["opac"].reshape(-1), 1e-4, "dL_dopacity", frac=0.995)
