"""Closed-form known-answer tests for the CPU oracle.

No golden vectors exist in the reference; these pin the oracle to the maths of the
credited algorithm (Appendix B) on cases whose answer can be written down by hand.
"""
import math

import numpy as np

from divshot_b200.scenes import look_at_camera
from oracle import oracle as orc
from util import orc_cam

W = H = 64


def _cam(bg=(0, 0, 0)):
    return look_at_camera((0, 0, 0), (0, 0, 1), W, H, bg=bg)


def _one(pos, s, o, rgb_dc, cam, deg=0, shN=None, quat=(1, 0, 0, 0)):
    n = len(pos)
    means = np.asarray(pos, np.float32).reshape(n, 3)
    scales = np.broadcast_to(np.asarray(s, np.float32), (n, 3)).copy()
    quats = np.broadcast_to(np.asarray(quat, np.float32), (n, 4)).copy()
    opac = np.broadcast_to(np.asarray(o, np.float32), (n,)).copy()
    sh0 = ((np.asarray(rgb_dc, np.float32).reshape(-1, 3) - 0.5) / 0.28209479177387814).astype(np.float32)
    sh0 = np.broadcast_to(sh0, (n, 3)).copy()
    K = (deg + 1) ** 2
    shN = np.zeros((n, K - 1, 3), np.float32) if shN is None else shN
    oc = orc_cam(cam, deg, flags=orc.FLAG_INPUT_ACTIVATED)
    return oc, (means, scales, quats, opac, sh0, shN)


def test_expf_accuracy_and_determinism():
    xs = np.concatenate([np.linspace(-87, 88, 20001), np.random.default_rng(0).normal(0, 3, 20000)]).astype(np.float32)
    got = np.array([orc.expf(float(x)) for x in xs], np.float32)
    ref = np.exp(xs.astype(np.float64))
    ulp = np.abs(got.astype(np.float64) - ref) / np.spacing(ref.astype(np.float32)).astype(np.float64)
    assert ulp.max() <= 2.0, ulp.max()
    assert orc.expf(0.0) == 1.0


def test_g1_single_isotropic_gaussian_closed_form():
    cam = _cam()
    z0, s, o = 4.0, 0.05, 0.8
    oc, arr = _one([(0, 0, z0)], s, o, (0.9, 0.5, 0.2), cam)
    f = orc.forward(oc, *arr, threads=1)
    fx = W / (2 * cam.tanfovx)
    a = (fx * s / z0) ** 2 + 0.3
    assert abs(f.conic_opacity[0, 0] - 1 / a) < 1e-5 / a and abs(f.conic_opacity[0, 1]) < 1e-7
    assert f.radii[0] == math.ceil(3 * math.sqrt(a + math.sqrt(0.1)))
    mu = (W - 1) / 2
    assert np.allclose(f.mean2D[0], [mu, mu], atol=1e-5)
    r = f.radii[0]
    lo, hi = int((mu - r) / 16), int((mu + r + 15) / 16)
    assert list(f.rect[0]) == [lo, lo, hi, hi] and f.tiles_touched[0] == (hi - lo) ** 2
    ys, xs = np.mgrid[0:H, 0:W]
    alpha = o * np.exp(-((xs - mu) ** 2 + (ys - mu) ** 2) / (2 * a))
    alpha = np.where(alpha < 1 / 255, 0, np.minimum(alpha, 0.99))
    tiles_ok = (xs // 16 >= lo) & (xs // 16 < hi) & (ys // 16 >= lo) & (ys // 16 < hi)
    alpha = alpha * tiles_ok
    for ch, c in enumerate((0.9, 0.5, 0.2)):
        assert np.allclose(f.image[ch], c * alpha, atol=2e-6)
    assert np.allclose(f.final_T.reshape(H, W), 1 - alpha, atol=2e-6)
    assert ((f.n_contrib.reshape(H, W) == 1) == (alpha > 0)).all()


def test_g2_equal_depth_tie_breaks_by_index():
    cam = _cam()
    oc, arr = _one([(0, 0, 4.0), (0, 0, 4.0)], 0.05, 0.5, [(1, 0, 0), (0, 1, 0)], cam)
    f = orc.forward(oc, *arr, threads=1)
    r0, r1 = f.ranges[f.rect[0, 1] * 4 + f.rect[0, 0]]
    assert list(f.point_list[r0:r1]) == [0, 1]
    c = int((W - 1) / 2)
    a = f.conic_opacity[0, 3] * math.exp(-0.5 * f.conic_opacity[0, 0] * 2 * (0.5 ** 2))
    assert abs(f.image[0, c, c] - a) < 1e-5 and abs(f.image[1, c, c] - a * (1 - a)) < 1e-5
    # every tile list is sorted by (depth bits, index)
    for t in range(f.ranges.shape[0]):
        ids = f.point_list[f.ranges[t, 0]:f.ranges[t, 1]].astype(np.int64)
        k = f.depth[ids].view(np.uint32).astype(np.int64) * (1 << 24) + ids
        assert (np.diff(k) > 0).all()


def test_g3_near_cull_boundary():
    cam = _cam()
    zs = [0.2, np.nextafter(np.float32(0.2), np.float32(1)), 0.1, -1.0]
    oc, arr = _one([(0, 0, z) for z in zs], 0.001, 0.5, (0.5, 0.5, 0.5), cam)
    f = orc.forward(oc, *arr, threads=1)
    assert list(f.radii > 0) == [False, True, False, False]


def test_g4_off_axis_beyond_the_ewa_clamp_closed_form():
    """A splat at x/z = 1.6 tanfov: the EWA Jacobian is evaluated at the clamped ray 1.3 tanfov (in-tree restatement:
    gsplat_intersect.hlsl:88-94) while mean2D is the true projection; the gradient through the clamped coordinate is
    zero (Appendix B.5), so dL/dmean.x comes from the mean2D path alone."""
    cam = _cam()
    z0, s, o = 4.0, 0.6, 0.7
    x0 = 1.6 * cam.tanfovx * z0
    oc, arr = _one([(x0, 0, z0)], s, o, (0.9, 0.5, 0.2), cam)
    f = orc.forward(oc, *arr, threads=1)
    fx = W / (2 * cam.tanfovx)
    fy = H / (2 * cam.tanfovy)
    tx = 1.3 * cam.tanfovx * z0  # clamped
    cxx = s * s * ((fx / z0) ** 2 + (fx * tx / z0 ** 2) ** 2) + 0.3
    cyy = s * s * (fy / z0) ** 2 + 0.3
    assert abs(f.conic_opacity[0, 0] - 1 / cxx) < 2e-5 / cxx and abs(f.conic_opacity[0, 2] - 1 / cyy) < 2e-5 / cyy
    assert abs(f.conic_opacity[0, 1]) < 1e-7
    mid = 0.5 * (cxx + cyy)
    assert f.radii[0] == math.ceil(3 * math.sqrt(mid + math.sqrt(max(0.1, mid * mid - cxx * cyy))))
    # mean2D is NOT clamped: ndc = x / (z tanfov) = 1.6 -> pixel ((1.6 + 1) W - 1) / 2, right of the image
    assert abs(f.mean2D[0, 0] - ((1.6 + 1) * W - 1) / 2) < 1e-3 and f.mean2D[0, 0] > W
    assert f.tiles_touched[0] > 0 and f.image.max() > 0.01  # its footprint still reaches the screen
    g = np.random.default_rng(0).normal(size=(3, H, W)).astype(np.float32)
    b = orc.backward(oc, f, *arr, g, threads=1)
    # dL_dmean2D is w.r.t. the ndc coordinate (the 0.5 W pixel factor already applied): d ndc.x / dx = 1 / (tanfov z)
    k = 1.0 / (cam.tanfovx * z0)
    assert abs(b.dL_dmeans3D[0, 0] - b.dL_dmean2D[0, 0] * k) <= 2e-5 * abs(b.dL_dmeans3D[0, 0]) + 1e-9
    # an unclamped twin (x/z = 1.2 tanfov) for contrast: there the covariance path does contribute to dL/dx
    oc2, arr2 = _one([(1.2 * cam.tanfovx * z0, 0, z0)], s, o, (0.9, 0.5, 0.2), cam)
    f2 = orc.forward(oc2, *arr2, threads=1)
    b2 = orc.backward(oc2, f2, *arr2, g, threads=1)
    assert abs(b2.dL_dmeans3D[0, 0] - b2.dL_dmean2D[0, 0] * k) > 1e-3 * abs(b2.dL_dmeans3D[0, 0])


def test_g5_opaque_stack_early_out():
    cam = _cam()
    n = 12
    oc, arr = _one([(0, 0, 3.0 + 0.1 * i) for i in range(n)], 0.2, 0.9, (0.5, 0.5, 0.5), cam)
    f = orc.forward(oc, *arr, threads=1)
    c = int((W - 1) / 2)
    # walk the centre pixel in fp32 exactly as B.3 prescribes
    T = np.float32(1); last = 0
    for i in range(n):
        d = np.float32(f.mean2D[i, 0] - c)
        p = np.float32(-0.5) * (f.conic_opacity[i, 0] * d * d + f.conic_opacity[i, 2] * d * d) - f.conic_opacity[i, 1] * d * d
        al = min(np.float32(0.99), np.float32(f.conic_opacity[i, 3] * np.exp(np.float32(p), dtype=np.float32)))
        if al < np.float32(1 / 255):
            continue
        Tn = np.float32(T * (np.float32(1) - al))
        if Tn < np.float32(1e-4):
            break
        T = Tn; last = i + 1
    assert 0 < last < n
    assert f.n_contrib.reshape(H, W)[c, c] == last
    assert abs(f.final_T.reshape(H, W)[c, c] - T) < 1e-9


def test_g6_sh_colour_and_clamp_mask():
    cam = _cam()
    rng = np.random.default_rng(5)
    n, deg = 64, 3
    pos = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(2, 6, n)], 1)
    shN = rng.normal(0, 0.6, (n, 15, 3)).astype(np.float32)
    oc, arr = _one(pos, 0.02, 0.5, rng.uniform(0, 1, (n, 3)), cam, deg=deg, shN=shN)
    f = orc.forward(oc, *arr, threads=1)
    import autograd_ref as ar
    import torch
    d = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    bas = ar.sh_basis(deg, torch.tensor(d)).numpy()
    coeffs = np.concatenate([arr[4][:, None, :], shN], 1).astype(np.float64)
    col = (bas[:, :, None] * coeffs).sum(1) + 0.5
    vis = f.radii > 0
    assert vis.sum() > 30 and (col[vis] < 0).any()
    assert np.allclose(f.rgb[vis], np.maximum(col[vis], 0), atol=3e-6)
    sure = np.abs(col) > 1e-5
    assert ((f.clamped.astype(bool) == (col < 0)) | ~sure)[vis].all()


def test_parallel_tile_sort_equals_literal_radix_sort():
    """orc_bin_sort (counting sort by tile + per-tile sorts, parallel) vs the literal single stable radix sort."""
    import ctypes as C

    from divshot_b200.scenes import make_scene
    from util import scene_arrays
    sc = make_scene(N=20000, width=160, height=128, sh_degree=0, seed=77)
    sc.means3D[:, 2] = np.round(sc.means3D[:, 2] * 4) / 4  # plenty of equal depths -> index tie-breaks
    sc.log_scales += 0.6
    oc = orc_cam(sc.cameras[0], 0)
    f = orc.forward(oc, *scene_arrays(sc), render=False)
    # rebuild the unsorted keys exactly as A3 emits them, then sort them the literal way
    L = orc.lib()
    gx = (160 + 15) // 16; T = gx * ((128 + 15) // 16)
    keys = []; vals = []
    for i in np.nonzero(f.radii > 0)[0]:
        x0, y0, x1, y1 = f.rect[i]
        for y in range(y0, y1):
            for x in range(x0, x1):
                keys.append(((y * gx + x) << 32) | int(f.depth[i:i + 1].view(np.uint32)[0])); vals.append(i)
    keys = np.array(keys, np.uint64); vals = np.array(vals, np.uint32); rng = np.zeros((T, 2), np.uint32)
    L.orc_radix_check(C.c_int32(T), C.c_int64(len(keys)), keys.ctypes.data_as(C.c_void_p),
                      vals.ctypes.data_as(C.c_void_p), rng.ctypes.data_as(C.c_void_p))
    assert len(keys) == f.D and np.array_equal(vals, f.point_list) and np.array_equal(keys, f.keys)
    assert np.array_equal(rng, f.ranges)
