"""Metamorphic properties of the ORACLE: relations any correct implementation of the credited algorithm
must satisfy, whatever its internals.  They pin the oracle — the anchor of every parity claim for the compositing stages,
for which the reference holds no code — from a side the closed-form KATs and the float64 autograd cross-check do not:
invariance under relabelling and under rigid motion, the affine role of the background, the neutrality of transparent
and of never-visible splats, linearity of the backward in dL/dpixel."""
import copy
import dataclasses

import numpy as np
import pytest

from divshot_b200.scenes import _flat, make_scene
from oracle import oracle as orc
from util import assert_close, orc_cam, rel_err, scene_arrays


def _scene(seed=3, deg=2, N=1500):
    sc = make_scene(N=N, width=80, height=56, sh_degree=deg, seed=seed, normalise_quats=True)
    sc.log_scales += 1.0
    return sc


def _render(sc, cam=None, deg=None):
    cam = cam or sc.cameras[0]
    deg = sc.sh_degree if deg is None else deg
    oc = orc_cam(cam, deg)
    return oc, orc.forward(oc, *scene_arrays(sc), threads=1)


def test_relabelling_the_gaussians_changes_nothing_visible():
    sc = _scene()
    sc.means3D[:, 2] += np.linspace(0, 1e-3, sc.N, dtype=np.float32)  # no exact depth ties (those break by index, by design)
    _, f = _render(sc)
    perm = np.random.default_rng(0).permutation(sc.N)
    sp = dataclasses.replace(sc, means3D=sc.means3D[perm], log_scales=sc.log_scales[perm], quats=sc.quats[perm],
                             logit_opac=sc.logit_opac[perm], sh0=sc.sh0[perm], shN=sc.shN[perm])
    oc, fp = _render(sp)
    assert np.array_equal(fp.image, f.image) and np.array_equal(fp.final_T, f.final_T)
    assert np.array_equal(fp.radii, f.radii[perm]) and fp.D == f.D
    inv = np.argsort(perm)  # old id -> new id
    for t in range(f.ranges.shape[0]):  # the same splats, in the same depth order, under their new names
        a = f.point_list[f.ranges[t, 0]:f.ranges[t, 1]]
        b = fp.point_list[fp.ranges[t, 0]:fp.ranges[t, 1]]
        assert np.array_equal(inv[a], b)
    b1 = orc.backward(oc, fp, *scene_arrays(sp), sc.dL_dpix[0], threads=1)
    b0 = orc.backward(orc_cam(sc.cameras[0], sc.sh_degree), f, *scene_arrays(sc), sc.dL_dpix[0], threads=1)
    assert np.array_equal(b1.dL_dmeans3D, b0.dL_dmeans3D[perm]) and np.array_equal(b1.dL_dshN, b0.dL_dshN[perm])


def _quat_mul(a, b):
    r1, x1, y1, z1 = a; r2, x2, y2, z2 = b.T
    return np.stack([r1 * r2 - x1 * x2 - y1 * y2 - z1 * z2, r1 * x2 + x1 * r2 + y1 * z2 - z1 * y2,
                     r1 * y2 - x1 * z2 + y1 * r2 + z1 * x2, r1 * z2 + x1 * y2 - y1 * x2 + z1 * r2], 1)


def test_moving_scene_and_camera_together_leaves_the_image_alone():
    """Degree 0 (no view-dependent colour to rotate): x -> R x + t on the means, q -> q_R q on the rotations, and the camera's
    world->view matrix composed with the inverse motion."""
    sc = _scene(seed=5, deg=0)
    _, f = _render(sc)
    ang = 0.7
    axis = np.array([0.3, -0.5, 0.8]); axis /= np.linalg.norm(axis)
    qR = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])
    r, x, y, z = qR
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                  [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                  [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])
    t = np.array([0.4, -1.1, 2.0])
    moved = dataclasses.replace(sc, means3D=(sc.means3D.astype(np.float64) @ R.T + t).astype(np.float32),
                                quats=_quat_mul(qR, sc.quats.astype(np.float64)).astype(np.float32))
    cam = sc.cameras[0]
    V = np.asarray(cam.view, np.float64).reshape(4, 4).T
    PV = np.asarray(cam.proj, np.float64).reshape(4, 4).T
    M = np.eye(4); M[:3, :3] = R; M[:3, 3] = t
    Minv = np.linalg.inv(M)
    cam2 = dataclasses.replace(cam, view=_flat(V @ Minv), proj=_flat(PV @ Minv),
                               campos=(R @ np.asarray(cam.campos, np.float64) + t).astype(np.float32))
    _, f2 = _render(moved, cam2)
    assert (f2.radii > 0).sum() >= 0.98 * (f.radii > 0).sum()
    assert rel_err(f2.image, f.image) < 2e-3  # fp32 re-association of the projection chain moves a few 1/255 decisions
    assert np.abs(f2.image - f.image).mean() < 2e-5


def test_background_enters_affinely_through_the_final_transmittance():
    sc = _scene(seed=7)
    _, f0 = _render(sc)
    bg = np.array([0.25, 0.5, 1.0], np.float32)
    cam = dataclasses.replace(sc.cameras[0], bg=bg)
    _, f1 = _render(sc, cam)
    T = f0.final_T.reshape(cam.height, cam.width)
    assert np.array_equal(f1.final_T, f0.final_T) and np.array_equal(f1.n_contrib, f0.n_contrib)
    assert np.allclose(f1.image, f0.image + bg[:, None, None] * T[None], atol=1e-6)


def test_transparent_and_never_visible_splats_are_neutral():
    sc = _scene(seed=9)
    oc, f = _render(sc)
    extra = 200
    rng = np.random.default_rng(1)
    more = dataclasses.replace(
        sc, means3D=np.concatenate([sc.means3D, np.concatenate([sc.means3D[:extra // 2], -np.abs(rng.normal(5, 1, (extra // 2, 3)))]).astype(np.float32)]),
        log_scales=np.concatenate([sc.log_scales, sc.log_scales[:extra]]), quats=np.concatenate([sc.quats, sc.quats[:extra]]),
        logit_opac=np.concatenate([sc.logit_opac, np.concatenate([np.full(extra // 2, -12.0), sc.logit_opac[:extra // 2]]).astype(np.float32)]),
        sh0=np.concatenate([sc.sh0, sc.sh0[:extra]]), shN=np.concatenate([sc.shN, sc.shN[:extra]]))
    oc2, f2 = _render(more)  # first half: opacity 6e-6 (below 1/255 everywhere); second half: behind the camera
    assert np.array_equal(f2.image, f.image) and np.array_equal(f2.final_T, f.final_T)
    assert (f2.radii[sc.N + extra // 2:] == 0).all()
    b = orc.backward(oc2, f2, *scene_arrays(more), sc.dL_dpix[0], threads=1)
    b0 = orc.backward(oc, f, *scene_arrays(sc), sc.dL_dpix[0], threads=1)
    assert not b.dL_dmeans3D[sc.N:].any() and not b.dL_dshN[sc.N:].any() and not b.dL_dopacities[sc.N:].any()
    assert np.array_equal(b.dL_dmeans3D[:sc.N], b0.dL_dmeans3D) and np.array_equal(b.dL_dquats[:sc.N], b0.dL_dquats)


@pytest.mark.parametrize("deg", [0, 3])
def test_backward_is_linear_in_the_pixel_gradient(deg):
    sc = _scene(seed=11, deg=deg)
    oc, f = _render(sc)
    rng = np.random.default_rng(2)
    g1 = sc.dL_dpix[0]
    g2 = rng.normal(size=g1.shape).astype(np.float32)
    a, b_ = 1.5, -0.25
    arrays = scene_arrays(sc)
    B1 = orc.backward(oc, f, *arrays, g1, threads=1)
    B2 = orc.backward(oc, f, *arrays, g2, threads=1)
    B12 = orc.backward(oc, f, *arrays, (a * g1 + b_ * g2).astype(np.float32), threads=1)
    for k in ("dL_dmeans3D", "dL_dscales", "dL_dquats", "dL_dopacities", "dL_dsh0", "dL_dshN"):
        want = a * getattr(B1, k).astype(np.float64) + b_ * getattr(B2, k).astype(np.float64)
        if want.size:
            assert_close(getattr(B12, k), want, 2e-4, k)
