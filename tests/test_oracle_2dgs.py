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


def _cull_ellipse(T, o, cx, cy):
    """numpy (float32) restatement of the sub-tile cull ellipse of csrc/preprocess_fwd.cu (surfel_preprocess_fwd_kernel): centre and
    conic (a, b, c) of the set emission tests the 8x4-pixel boxes against (a dx^2 + b dx dy + c dy^2 <= ~1), or None = no culling."""
    f = np.float32
    m2 = f(2.0) * np.log(f(255.0) * o) * f(1.0001) + f(1e-3)
    if not m2 > 0:
        return "empty"
    dist_k = m2 * (T[6] * T[6] + T[7] * T[7]) - T[8] * T[8]
    if not dist_k < f(-1e-6) * T[8] * T[8]:
        return None
    u = T[0:3] - cx * T[6:9]; v = T[3:6] - cy * T[6:9]; w = T[6:9]
    g0 = m2 / dist_k; g2 = f(-1.0) / dist_k
    kx = g0 * (u[0] * w[0] + u[1] * w[1]) + g2 * u[2] * w[2]
    ky = g0 * (v[0] * w[0] + v[1] * w[1]) + g2 * v[2] * w[2]
    hx = max(kx * kx - (g0 * (u[0] * u[0] + u[1] * u[1]) + g2 * u[2] * u[2]), f(0))
    hy = max(ky * ky - (g0 * (v[0] * v[0] + v[1] * v[1]) + g2 * v[2] * v[2]), f(0))
    hxy = kx * ky - (g0 * (u[0] * v[0] + u[1] * v[1]) + g2 * u[2] * v[2])
    lim = np.sqrt(hx * hy); hxy = min(max(hxy, -lim), lim)
    rf = np.sqrt(f(0.5) * m2)
    r = rf + np.sqrt(kx * kx + ky * ky) + f(0.25)
    sxx = hx * f(1.002) + r * r; syy = hy * f(1.002) + r * r; sxy = hxy * f(1.002)
    det = sxx * syy - sxy * sxy
    exk = np.sqrt(max(f(1e-4), hx)); eyk = np.sqrt(max(f(1e-4), hy))
    xlo, xhi = min(kx - exk, -rf), max(kx + exk, rf)
    ylo, yhi = min(ky - eyk, -rf), max(ky + eyk, rf)
    sx = f(1.4143) * (f(0.5) * (xhi - xlo) * f(1.001) + f(0.05)); sy = f(1.4143) * (f(0.5) * (yhi - ylo) * f(1.001) + f(0.05))
    if sxx < 1e12 and syy < 1e12 and det > 0 and det < sx * sx * sy * sy:
        return (cx + kx, cy + ky, syy / det, f(-2.0) * sxy / det, sxx / det)
    if sx < 1e6 and sy < 1e6:
        return (cx + f(0.5) * (xlo + xhi), cy + f(0.5) * (ylo + yhi), f(1.0) / (sx * sx), f(0.0), f(1.0) / (sy * sy))
    return None


def test_cull_ellipse_contains_every_pixel_the_oracle_blends():
    """The surfel cull ellipse (exact projected alpha >= 1/255 ellipse from the dual conic, inflated to contain the low-pass disk;
    DESIGN.md section 8) is conservative: every (pixel, surfel) pair that passes the oracle's S.2 tests lies inside it.  The
    kernel's construction is restated in numpy above; the GPU tier checks the kernel's actual mask bits the same way."""
    total = outside = culled_all = 0
    for seed, (N, W, H, ls) in enumerate([(700, 96, 64, 1.2), (500, 128, 80, 0.3), (300, 160, 100, 2.0)]):
        sc = make_scene(N=N, width=W, height=H, sh_degree=0, seed=431 + seed, normalise_quats=False)
        sc.log_scales += ls
        sc.logit_opac[::3] += 3.0
        sc.logit_opac[1::7] -= 6.0
        f = orc.forward2d(orc_cam(sc.cameras[0], 0), *scene_arrays(sc), render=False)
        ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
        for g in np.nonzero(np.asarray(f.radii) > 0)[0]:
            T = f.transmat[g].astype(np.float32)
            o = np.float32(f.opacity[g]); cx, cy = f.mean2D[g].astype(np.float32)
            Td = T.astype(np.float64); Tu, Tv, Tw = Td[0:3], Td[3:6], Td[6:9]
            k = xs[..., None] * Tw - Tu; l = ys[..., None] * Tw - Tv
            pv = np.cross(k, l)
            with np.errstate(all="ignore"):
                uu, vv = pv[..., 0] / pv[..., 2], pv[..., 1] / pv[..., 2]
            rho3d = uu * uu + vv * vv
            rho2d = 2.0 * ((cx - xs) ** 2 + (cy - ys) ** 2)
            dep = np.where(rho3d <= rho2d, uu * Tw[0] + vv * Tw[1] + Tw[2], Tw[2])
            alpha = np.minimum(0.99, o * np.exp(-0.5 * np.minimum(rho3d, rho2d)))
            contrib = (pv[..., 2] != 0) & (dep >= 0.2) & (alpha >= 1 / 255)
            total += int(contrib.sum())
            e = _cull_ellipse(T, o, cx, cy)
            if e is None:
                continue
            if isinstance(e, str):  # opacity below 1/255: the mask is empty and nothing may blend
                outside += int(contrib.sum()); culled_all += 1
                continue
            ex, ey, a, b, c = (np.float64(t) for t in e)
            dx, dy = ex - xs, ey - ys
            outside += int((contrib & ~(a * dx * dx + b * dx * dy + c * dy * dy <= 1.0011)).sum())
    assert total > 200000 and culled_all > 10
    assert outside == 0


def test_recursive_halving_warp_sum_leaves_value_i_on_lanes_2i_and_2i_plus_1():
    """The backward's warp reduction (csrc/surfel.cu): 16 values per lane are summed over the 32 lanes with 8 + 4 + 2 + 1 + 1
    butterfly shuffles by keeping one half and trading the other at every level; lane L must end with the warp total of value
    L >> 1, which the four gather shuffles (source lane 8 (lane & 3) + 2 c) then turn into the float4 of lanes 0..3.  Emulated
    lane by lane in numpy with the kernel's own select / shuffle pattern."""
    rng = np.random.default_rng(5)
    g = rng.integers(-1000, 1000, (32, 16)).astype(np.float64)  # [lane][value]; integers: exact sums in any order
    lane = np.arange(32)

    def shfl_xor(v, m):
        return v[lane ^ m]

    def level(vals, bit, half):
        h = (lane & bit) != 0
        out = []
        for k in range(half):
            send = np.where(h, vals[k], vals[k + half]); keep = np.where(h, vals[k + half], vals[k])
            out.append(keep + shfl_xor(send, bit))
        return out

    a = level([g[:, k] for k in range(16)], 16, 8)
    b = level(a, 8, 4)
    c = level(b, 4, 2)
    t = level(c, 2, 1)[0]
    t = t + shfl_xor(t, 1)
    want = g.sum(axis=0)
    assert np.array_equal(t, want[lane >> 1])
    for ln in range(4):  # the gather for the four 128-bit reductions
        got = [t[8 * (ln & 3) + 2 * cc] for cc in range(4)]
        assert got == [want[4 * ln + cc] for cc in range(4)]
