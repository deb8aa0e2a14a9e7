"""Auxiliary outputs (depth / alpha maps) restated with the ORACLE's existing functions, by linearity:
the compositor is linear in the per-splat "colour", so depth = sum_k w_k z_k and alpha = sum_k w_k are the image of the
colour triple (z_k, 1, 0) over a zero background, and the gradient of <depth, gD> + <alpha, gA> is the compositor's
backward for that triple — its geometry part (dL/dmean2D, dL/dconic, dL/dopacity) ADDS to the colour loss's, its
"colour" part is dL/dz_k, which reaches the 3-D mean through row 2 of the view matrix.  The CUDA path
(divshot_b200/csrc/aux_outputs.cu + dvs_rast_forward_aux / dvs_rast_backward_aux) is the same composition of the same
kernels; tests/test_aux_outputs.py checks this restatement against float64 autograd and the CUDA path against it."""
import copy
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import oracle as orc


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cam_bg0(cam):
    c = copy.copy(cam)
    c.bg[:] = [0.0, 0.0, 0.0]
    return c


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def normal_ops():
    """Host build of divshot_b200/csrc/aux_normal_ops.h (the arithmetic the CUDA kernels call)."""
    src = os.path.join(ROOT, "tests", "native", "aux_normal_host.cpp")
    hdr = os.path.join(ROOT, "divshot_b200", "csrc", "aux_normal_ops.h")
    out = os.path.join(ROOT, "build", "test_aux_normal_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-I", os.path.dirname(hdr), src, "-o", out])
    L = C.CDLL(out)
    L.t_normals_forward.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_longlong] + [C.c_void_p] * 3
    L.t_normals_backward.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_longlong] + [C.c_void_p] * 2
    return L


def normals(cam, means3D, scales, quats, activated=False):
    """-> (n_v [N,3], axis [N], flip [N]): view-space normal of every Gaussian and the decisions behind it."""
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, np.float32))  # noqa: E731
    means3D, scales, quats = f32(means3D), f32(scales), f32(quats)
    N = means3D.shape[0]
    view = np.array(list(cam.view), np.float32)
    n_v = np.zeros((N, 3), np.float32); axis = np.zeros(N, np.int32); flip = np.zeros(N, np.float32)
    normal_ops().t_normals_forward(_p(quats), _p(scales), _p(means3D), _p(view), int(activated), N, _p(n_v), _p(axis), _p(flip))
    return n_v, axis, flip


def aux_colours(fwd):
    N = fwd.depth.shape[0]
    col = np.zeros((N, 3), np.float32)
    col[:, 0] = fwd.depth
    col[:, 1] = 1.0
    col[fwd.radii <= 0] = 0.0
    return col


def forward_aux(cam, fwd):
    """-> [2,H,W] float32 (depth, alpha) from the lists of an oracle forward."""
    L = orc.lib()
    W, H = cam.width, cam.height
    image = np.zeros((3, H, W), np.float32); final_T = np.zeros(H * W, np.float32)
    n_contrib = np.zeros(H * W, np.uint32); fragile = np.zeros(H * W, np.uint8)
    col = aux_colours(fwd)
    plist = fwd.point_list if fwd.D else np.zeros(1, np.uint32)
    c0 = _cam_bg0(cam)
    L.orc_render_fwd(C.byref(c0), _p(fwd.ranges), _p(plist), _p(fwd.mean2D), _p(fwd.conic_opacity), _p(col), _p(image),
                     _p(final_T), _p(n_contrib), _p(fragile), C.c_int32(1))
    assert np.array_equal(n_contrib, fwd.n_contrib), "the auxiliary pass walks the same lists to the same depth"
    return image[:2].copy()


def forward_normals(cam, fwd, means3D, scales, quats, activated=False):
    """-> [3,H,W] float32: normal map = sum_k w_k n_k from the lists of an oracle forward."""
    L = orc.lib()
    W, H = cam.width, cam.height
    image = np.zeros((3, H, W), np.float32); final_T = np.zeros(H * W, np.float32)
    n_contrib = np.zeros(H * W, np.uint32); fragile = np.zeros(H * W, np.uint8)
    n_v, _, _ = normals(cam, means3D, scales, quats, activated)
    n_v[fwd.radii <= 0] = 0
    plist = fwd.point_list if fwd.D else np.zeros(1, np.uint32)
    c0 = _cam_bg0(cam)
    L.orc_render_fwd(C.byref(c0), _p(fwd.ranges), _p(plist), _p(fwd.mean2D), _p(fwd.conic_opacity), _p(n_v), _p(image),
                     _p(final_T), _p(n_contrib), _p(fragile), C.c_int32(1))
    return image


def backward_with_aux(cam, fwd, means3D, scales, quats, opacities, sh0, shN, dL_dpix, dL_daux, dL_dnormal=None,
                      activated=False):
    """Gradients of <image, dL_dpix> + <depth, dL_daux[0]> + <alpha, dL_daux[1]> (+ <normal map, dL_dnormal>) w.r.t. the
    stored parameters."""
    L = orc.lib()
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, np.float32))  # noqa: E731
    means3D, scales, quats, sh0 = f32(means3D), f32(scales), f32(quats), f32(sh0)
    opacities = f32(opacities).reshape(-1)
    N = means3D.shape[0]
    KR = cam.sh_rest_alloc
    shN = f32(shN).reshape(N, KR, 3) if KR > 0 else np.zeros((N, 0, 3), np.float32)
    W, H = cam.width, cam.height
    plist = fwd.point_list if fwd.D else np.zeros(1, np.uint32)

    def render_bwd(c, colours, dpix):
        g_m2 = np.empty((N, 2), np.float32); g_abs = np.empty((N, 2), np.float32); g_con = np.empty((N, 3), np.float32)
        g_op = np.empty(N, np.float32); g_col = np.empty((N, 3), np.float32)
        L.orc_render_bwd(C.byref(c), C.c_int32(N), _p(fwd.ranges), _p(plist), _p(fwd.mean2D), _p(fwd.conic_opacity),
                         _p(f32(colours)), _p(fwd.final_T), _p(fwd.n_contrib), _p(f32(dpix).reshape(3, H, W)), _p(g_m2),
                         _p(g_abs), _p(g_con), _p(g_op), _p(g_col), C.c_int32(1))
        return g_m2, g_con, g_op, g_col

    m2a, cona, opa, cola = render_bwd(cam, fwd.rgb, dL_dpix)
    daux3 = np.zeros((3, H, W), np.float32); daux3[:2] = np.asarray(dL_daux, np.float32).reshape(2, H, W)
    m2b, conb, opb, colb = render_bwd(_cam_bg0(cam), aux_colours(fwd), daux3)
    g_m2, g_con, g_op = m2a + m2b, cona + conb, opa + opb
    dq_normal = None
    if dL_dnormal is not None:  # third pass: colour = view-space normal; its colour sums are dL/dn_k
        n_v, axis, flip = normals(cam, means3D, scales, quats, activated)
        n_v[fwd.radii <= 0] = 0
        m2c, conc, opc, colc = render_bwd(_cam_bg0(cam), n_v, np.asarray(dL_dnormal, np.float32).reshape(3, H, W))
        g_m2, g_con, g_op = g_m2 + m2c, g_con + conc, g_op + opc
        dn = f32(colc * (fwd.radii > 0)[:, None])
        dq_normal = np.zeros((N, 4), np.float32)
        view = np.array(list(cam.view), np.float32)
        normal_ops().t_normals_backward(_p(quats), _p(axis), _p(flip), _p(view), int(activated), N, _p(dn), _p(dq_normal))
    d_means = np.empty((N, 3), np.float32); d_scales = np.empty((N, 3), np.float32); d_quats = np.empty((N, 4), np.float32)
    d_opac = np.empty(N, np.float32); d_sh0 = np.empty((N, 3), np.float32); d_shN = np.zeros((N, max(KR, 0), 3), np.float32)
    L.orc_preprocess_bwd(C.byref(cam), C.c_int32(N), _p(means3D), _p(scales), _p(quats), _p(opacities), _p(sh0), _p(shN),
                         _p(fwd.radii), _p(fwd.clamped), _p(g_m2), _p(g_con), _p(g_op), _p(cola), _p(d_means), _p(d_scales),
                         _p(d_quats), _p(d_opac), _p(d_sh0), _p(d_shN if KR > 0 else np.zeros(1, np.float32)))
    dz = colb[:, 0] * (fwd.radii > 0)
    row2 = np.array([cam.view[4 * c + 2] for c in range(3)], np.float32)  # d z_view / d mean
    d_means = d_means + dz[:, None] * row2[None, :]
    if dq_normal is not None:
        d_quats = d_quats + dq_normal
    return dict(means3D=d_means, scales=d_scales, quats=d_quats, opac=d_opac, sh0=d_sh0, shN=d_shN, dz=dz)
