"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE — see oracle/dvs_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  "Parity unpinned": the reference's own rasterizer is absent from the
reference tree; this restates the credited public algorithm.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLAG_INPUT_ACTIVATED = 1
FLAG_ANTIALIAS = 2


class OrcCamera(C.Structure):
    _fields_ = [
        ("view", C.c_float * 16),
        ("proj", C.c_float * 16),
        ("campos", C.c_float * 3),
        ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("bg", C.c_float * 3),
        ("scale_modifier", C.c_float),
        ("sh_degree", C.c_int32),
        ("sh_rest_alloc", C.c_int32),
        ("flags", C.c_int32),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdvs_oracle.so")
    src = os.path.join(_HERE, "dvs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_scan_tiles.restype = C.c_int64
        _LIB.orc_expf.restype = C.c_float
        _LIB.orc_expf.argtypes = [C.c_float]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_camera(view, proj, campos, tanfovx, tanfovy, width, height, bg=(0, 0, 0),
                scale_modifier=1.0, sh_degree=0, sh_rest_alloc=None, flags=0) -> OrcCamera:
    """view/proj: flat float32[16], element [4*c + r] = row r col c."""
    cam = OrcCamera()
    cam.view[:] = [float(x) for x in np.asarray(view, np.float32).reshape(16)]
    cam.proj[:] = [float(x) for x in np.asarray(proj, np.float32).reshape(16)]
    cam.campos[:] = [float(x) for x in np.asarray(campos, np.float32).reshape(3)]
    cam.tanfovx, cam.tanfovy = float(np.float32(tanfovx)), float(np.float32(tanfovy))
    cam.width, cam.height = int(width), int(height)
    cam.bg[:] = [float(x) for x in np.asarray(bg, np.float32).reshape(3)]
    cam.scale_modifier = float(scale_modifier)
    cam.sh_degree = int(sh_degree)
    K = (sh_degree + 1) ** 2
    cam.sh_rest_alloc = int(K - 1 if sh_rest_alloc is None else sh_rest_alloc)
    cam.flags = int(flags)
    return cam


@dataclass
class Forward:
    depth: np.ndarray
    radii: np.ndarray
    mean2D: np.ndarray
    cov3D: np.ndarray
    conic_opacity: np.ndarray
    rgb: np.ndarray
    clamped: np.ndarray
    tiles_touched: np.ndarray
    rect: np.ndarray
    point_offsets: np.ndarray
    D: int
    keys: np.ndarray
    point_list: np.ndarray
    ranges: np.ndarray
    image: np.ndarray
    final_T: np.ndarray
    n_contrib: np.ndarray
    fragile: np.ndarray
    extra: dict = field(default_factory=dict)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def forward(cam: OrcCamera, means3D, scales, quats, opacities, sh0, shN, threads: int = 0,
            render: bool = True) -> Forward:
    L = lib()
    means3D, scales, quats = _f32(means3D), _f32(scales), _f32(quats)
    opacities, sh0 = _f32(opacities).reshape(-1), _f32(sh0)
    N = means3D.shape[0]
    KR = cam.sh_rest_alloc
    shN = _f32(shN).reshape(N, KR, 3) if KR > 0 else np.zeros((N, 0, 3), np.float32)
    W, H = cam.width, cam.height
    gx, gy = (W + 15) // 16, (H + 15) // 16
    depth = np.empty(N, np.float32); radii = np.empty(N, np.int32)
    mean2D = np.empty((N, 2), np.float32); cov3D = np.empty((N, 6), np.float32)
    conic_opacity = np.empty((N, 4), np.float32); rgb = np.empty((N, 3), np.float32)
    clamped = np.empty((N, 3), np.uint8); tiles = np.empty(N, np.uint32)
    rect = np.empty((N, 4), np.int32)
    L.orc_preprocess_fwd(C.byref(cam), C.c_int32(N), _p(means3D), _p(scales), _p(quats), _p(opacities),
                         _p(sh0), _p(shN), _p(depth), _p(radii), _p(mean2D), _p(cov3D),
                         _p(conic_opacity), _p(rgb), _p(clamped), _p(tiles), _p(rect))
    offs = np.empty(N, np.uint32)
    D = int(L.orc_scan_tiles(C.c_int32(N), _p(tiles), _p(offs)))
    keys = np.empty(max(D, 1), np.uint64); plist = np.empty(max(D, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    L.orc_bin_sort(C.byref(cam), C.c_int32(N), _p(depth), _p(radii), _p(rect), _p(offs), C.c_int64(D),
                   _p(keys), _p(plist), _p(ranges))
    keys, plist = keys[:D], plist[:D]
    image = np.zeros((3, H, W), np.float32); final_T = np.zeros(H * W, np.float32)
    n_contrib = np.zeros(H * W, np.uint32); fragile = np.zeros(H * W, np.uint8)
    if render:
        L.orc_render_fwd(C.byref(cam), _p(ranges), _p(plist if D else np.zeros(1, np.uint32)), _p(mean2D),
                         _p(conic_opacity), _p(rgb), _p(image), _p(final_T), _p(n_contrib), _p(fragile),
                         C.c_int32(threads))
    return Forward(depth, radii, mean2D, cov3D, conic_opacity, rgb, clamped, tiles, rect, offs, D, keys,
                   plist, ranges, image, final_T, n_contrib, fragile)


@dataclass
class Backward:
    dL_dmean2D: np.ndarray
    dL_dmean2D_abs: np.ndarray
    dL_dconic: np.ndarray
    dL_dopacity_act: np.ndarray
    dL_dcolor: np.ndarray
    dL_dmeans3D: np.ndarray
    dL_dscales: np.ndarray
    dL_dquats: np.ndarray
    dL_dopacities: np.ndarray
    dL_dsh0: np.ndarray
    dL_dshN: np.ndarray


def backward(cam: OrcCamera, fwd: Forward, means3D, scales, quats, opacities, sh0, shN, dL_dpix,
             threads: int = 0) -> Backward:
    L = lib()
    means3D, scales, quats = _f32(means3D), _f32(scales), _f32(quats)
    opacities, sh0 = _f32(opacities).reshape(-1), _f32(sh0)
    N = means3D.shape[0]
    KR = cam.sh_rest_alloc
    shN = _f32(shN).reshape(N, KR, 3) if KR > 0 else np.zeros((N, 0, 3), np.float32)
    dL_dpix = _f32(dL_dpix).reshape(3, cam.height, cam.width)
    g_m2 = np.empty((N, 2), np.float32); g_abs = np.empty((N, 2), np.float32)
    g_con = np.empty((N, 3), np.float32); g_op = np.empty(N, np.float32); g_col = np.empty((N, 3), np.float32)
    plist = fwd.point_list if fwd.D else np.zeros(1, np.uint32)
    L.orc_render_bwd(C.byref(cam), C.c_int32(N), _p(fwd.ranges), _p(plist), _p(fwd.mean2D),
                     _p(fwd.conic_opacity), _p(fwd.rgb), _p(fwd.final_T), _p(fwd.n_contrib), _p(dL_dpix),
                     _p(g_m2), _p(g_abs), _p(g_con), _p(g_op), _p(g_col), C.c_int32(threads))
    d_means = np.empty((N, 3), np.float32); d_scales = np.empty((N, 3), np.float32)
    d_quats = np.empty((N, 4), np.float32); d_opac = np.empty(N, np.float32)
    d_sh0 = np.empty((N, 3), np.float32); d_shN = np.zeros((N, max(KR, 0), 3), np.float32)
    L.orc_preprocess_bwd(C.byref(cam), C.c_int32(N), _p(means3D), _p(scales), _p(quats), _p(opacities),
                         _p(sh0), _p(shN), _p(fwd.radii), _p(fwd.clamped), _p(g_m2), _p(g_con), _p(g_op),
                         _p(g_col), _p(d_means), _p(d_scales), _p(d_quats), _p(d_opac), _p(d_sh0),
                         _p(d_shN if KR > 0 else np.zeros(1, np.float32)))
    return Backward(g_m2, g_abs, g_con, g_op, g_col, d_means, d_scales, d_quats, d_opac, d_sh0, d_shN)


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int32(int(n)))


def expf(x: float) -> float:
    return float(lib().orc_expf(C.c_float(x)))


# ---- 2DGS ("surfel") variant: GaussianTrainConfig::modelType = 1 (spec S.1-S.4 in dvs_oracle.c) -------------------------
@dataclass
class Forward2D:
    depth: np.ndarray
    radii: np.ndarray
    mean2D: np.ndarray
    transmat: np.ndarray
    opacity: np.ndarray
    rgb: np.ndarray
    clamped: np.ndarray
    tiles_touched: np.ndarray
    rect: np.ndarray
    D: int
    point_list: np.ndarray
    ranges: np.ndarray
    image: np.ndarray
    final_T: np.ndarray
    n_contrib: np.ndarray
    fragile: np.ndarray


def forward2d(cam: OrcCamera, means3D, scales, quats, opacities, sh0, shN, threads: int = 0, render: bool = True) -> Forward2D:
    L = lib()
    means3D, scales, quats = _f32(means3D), _f32(scales), _f32(quats)
    opacities, sh0 = _f32(opacities).reshape(-1), _f32(sh0)
    N = means3D.shape[0]
    KR = cam.sh_rest_alloc
    shN = _f32(shN).reshape(N, KR, 3) if KR > 0 else np.zeros((N, 0, 3), np.float32)
    W, H = cam.width, cam.height
    gx, gy = (W + 15) // 16, (H + 15) // 16
    depth = np.empty(N, np.float32); radii = np.empty(N, np.int32); mean2D = np.empty((N, 2), np.float32)
    tm = np.empty((N, 9), np.float32); op = np.empty(N, np.float32); rgb = np.empty((N, 3), np.float32)
    clamped = np.empty((N, 3), np.uint8); tiles = np.empty(N, np.uint32); rect = np.empty((N, 4), np.int32)
    L.orc2_preprocess_fwd(C.byref(cam), C.c_int32(N), _p(means3D), _p(scales), _p(quats), _p(opacities), _p(sh0), _p(shN),
                          _p(depth), _p(radii), _p(mean2D), _p(tm), _p(op), _p(rgb), _p(clamped), _p(tiles), _p(rect))
    offs = np.empty(N, np.uint32)
    D = int(L.orc_scan_tiles(C.c_int32(N), _p(tiles), _p(offs)))
    keys = np.empty(max(D, 1), np.uint64); plist = np.empty(max(D, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    L.orc_bin_sort(C.byref(cam), C.c_int32(N), _p(depth), _p(radii), _p(rect), _p(offs), C.c_int64(D), _p(keys), _p(plist), _p(ranges))
    plist = plist[:D]
    image = np.zeros((3, H, W), np.float32); final_T = np.zeros(H * W, np.float32)
    n_contrib = np.zeros(H * W, np.uint32); fragile = np.zeros(H * W, np.uint8)
    if render:
        L.orc2_render_fwd(C.byref(cam), _p(ranges), _p(plist if D else np.zeros(1, np.uint32)), _p(mean2D), _p(tm), _p(op), _p(rgb),
                          _p(image), _p(final_T), _p(n_contrib), _p(fragile), C.c_int32(threads))
    return Forward2D(depth, radii, mean2D, tm, op, rgb, clamped, tiles, rect, D, plist, ranges, image, final_T, n_contrib, fragile)


@dataclass
class Backward2D:
    dL_dT: np.ndarray
    dL_dmean2D: np.ndarray
    dL_dopacity_act: np.ndarray
    dL_dcolor: np.ndarray
    dL_dmeans3D: np.ndarray
    dL_dscales: np.ndarray
    dL_dquats: np.ndarray
    dL_dopacities: np.ndarray
    dL_dsh0: np.ndarray
    dL_dshN: np.ndarray


def backward2d(cam: OrcCamera, fwd: Forward2D, means3D, scales, quats, opacities, sh0, shN, dL_dpix, threads: int = 0) -> Backward2D:
    L = lib()
    means3D, scales, quats = _f32(means3D), _f32(scales), _f32(quats)
    opacities, sh0 = _f32(opacities).reshape(-1), _f32(sh0)
    N = means3D.shape[0]
    KR = cam.sh_rest_alloc
    shN = _f32(shN).reshape(N, KR, 3) if KR > 0 else np.zeros((N, 0, 3), np.float32)
    dL_dpix = _f32(dL_dpix).reshape(3, cam.height, cam.width)
    g_T = np.empty((N, 9), np.float32); g_m2 = np.empty((N, 2), np.float32); g_op = np.empty(N, np.float32)
    g_col = np.empty((N, 3), np.float32)
    plist = fwd.point_list if fwd.D else np.zeros(1, np.uint32)
    L.orc2_render_bwd(C.byref(cam), C.c_int32(N), _p(fwd.ranges), _p(plist), _p(fwd.mean2D), _p(fwd.transmat), _p(fwd.opacity),
                      _p(fwd.rgb), _p(fwd.final_T), _p(fwd.n_contrib), _p(dL_dpix), _p(g_T), _p(g_m2), _p(g_op), _p(g_col),
                      C.c_int32(threads))
    d_means = np.empty((N, 3), np.float32); d_scales = np.empty((N, 3), np.float32); d_quats = np.empty((N, 4), np.float32)
    d_opac = np.empty(N, np.float32); d_sh0 = np.empty((N, 3), np.float32); d_shN = np.zeros((N, max(KR, 0), 3), np.float32)
    L.orc2_preprocess_bwd(C.byref(cam), C.c_int32(N), _p(means3D), _p(scales), _p(quats), _p(opacities), _p(sh0), _p(shN),
                          _p(fwd.radii), _p(fwd.clamped), _p(g_T), _p(g_m2), _p(g_op), _p(g_col), _p(d_means), _p(d_scales),
                          _p(d_quats), _p(d_opac), _p(d_sh0), _p(d_shN if KR > 0 else np.zeros(1, np.float32)))
    return Backward2D(g_T, g_m2, g_op, g_col, d_means, d_scales, d_quats, d_opac, d_sh0, d_shN)
