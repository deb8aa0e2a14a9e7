"""Shared pieces of the viewer hand-off tests: the test model, the host build of
divshot_b200/csrc/viewer_pack_ops.h, and the reference's own quantiser (oracle/_ref/libviewerpack_ref.so) when present."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libviewerpack_ref.so")
KEYS = ("means", "scales", "quats", "opac", "sh0", "shN")


def make_model(N, seed, active_degree=3):
    """A trained-looking model plus every edge the quantiser has: half-precision subnormal and overflowing scales,
    saturated opacities, zero / tiny / huge quaternions, all-zero SH rows, a negative c[0] of largest magnitude (the
    reference's signed-max quirk, values leave [-1, 1]), exact ties of the half rounding, signed zeros."""
    rng = np.random.default_rng(seed)
    m = dict(means=rng.normal(0, 3, (N, 3)), scales=rng.normal(-4, 1.2, (N, 3)), quats=rng.normal(0, 1, (N, 4)),
             opac=rng.normal(0, 2.5, N), sh0=rng.normal(0, 1.2, (N, 3)), shN=rng.normal(0, 0.2, (N, 45)))
    m = {k: v.astype(np.float32) for k, v in m.items()}
    if active_degree < 3:
        m["shN"][:, 3 * ((active_degree + 1) ** 2 - 1):] = 0
    if N >= 64:
        m["scales"][0] = [-20.0, -17.3, -11.0]        # below / inside the half subnormal range
        m["scales"][1] = [11.2, 12.0, 88.0]            # 65504 is the largest half; exp(88) ~ 1.6e38
        m["scales"][2] = [-104.0, 89.0, 0.0]           # float underflow / overflow of exp itself
        m["opac"][3:7] = [-40.0, 40.0, -104.0, 17.0]
        m["quats"][7] = [1e-20, 0, 0, 0]               # len2 underflows to 0: 0/0 and x/0
        m["quats"][8] = [3e19, 1e19, -2e19, 0.5]       # len2 overflows to inf
        m["quats"][9] = [0, 0, 0, 0]
        m["quats"][10] = [1, 0, 0, 0]
        m["shN"][11] = 0
        m["shN"][12] = 0; m["shN"][12, 0] = -0.0
        m["shN"][13, 0] = -5.0                         # signed-max quirk: c[0]/max < -1
        m["shN"][14, 0] = 5.0
        m["shN"][15] = 0; m["shN"][15, 7] = -1e-30
        m["shN"][16] = -np.abs(m["shN"][16]); m["shN"][16, 0] = -1e-3
        m["shN"][22] = 1e-30; m["shN"][22, 0] = -5.0   # quirk taken to the limit: c[0]/max = -5e30, beyond int64 after scaling
        m["shN"][23] = 0; m["shN"][23, 0] = -2.0       # scale stays +0 after the loop?  no: max(-2, |0|) = 0 -> no division
        # rows around the guards of the kernel's short-form division / float quantiser (viewer_pack_ops.h: pack_sh_rest_from)
        m["shN"][24] = m["shN"][24] * np.float32(2.0 ** -58)           # whole row tiny but inside the short form's range
        m["shN"][25, 5] = np.float32(2.0 ** -63)                       # one magnitude below it: the guarded general form
        m["shN"][26] = np.where(np.arange(45) % 2 == 0, 0.37, -0.37)   # every quotient exactly +-1
        m["shN"][27] = (2.0 * np.arange(45) * 45 / 2047.0 - 1.0); m["shN"][27, 44] = 1.0  # scale 1, values on k / S boundaries
        m["shN"][28] = m["shN"][28] * np.float32(2.0 ** 58)            # large row, still inside
        m["shN"][29] = m["shN"][29] * np.float32(2.0 ** 70)            # scale above the range: general form
        m["shN"][30] = 0; m["shN"][30, 9] = 0.25; m["shN"][30, 10] = -0.0  # zeros of both signs beside one non-zero
        m["shN"][31, 3] = 1e-40; m["shN"][31, 4] = -1e-37; m["shN"][31, 5] = -1e-45   # subnormal / tiny numerators, ordinary scale
        m["shN"][32, 0] = -np.abs(m["shN"][32]).max() * np.float32(1.0000001)         # c[0] negative, just past the scale
        m["sh0"][17] = [-1.7724539, 1.7724539, 0.0]    # colours at 0 / 1 / 0.5
        m["sh0"][18] = [1e-12, -1e-12, 300.0]
        m["means"][19] = [-0.0, 0.0, 1e-42]            # signed zero, float subnormal
        m["means"][20] = [1e30, -1e30, 7.0]
        # exact ties of the float -> half rounding (mantissa bit 12 set, lower bits clear) straight through the colour path
        tie = np.array([0x3C001000, 0x3C003000, 0xBC001000], np.uint32).view(np.float32)  # 1 + 2^-11, 1 + 3*2^-11
        m["sh0"][21] = (tie.astype(np.float64) - 0.5) / 0.28209479177387814
    return m


def host_ops():
    src = os.path.join(ROOT, "tests", "native", "viewer_pack_host.cpp")
    hdr = os.path.join(ROOT, "divshot_b200", "csrc", "viewer_pack_ops.h")
    out = os.path.join(ROOT, "build", "test_viewer_pack_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra",
                               "-I", os.path.dirname(hdr), src, "-o", out])
    L = C.CDLL(out)
    L.t_viewer_pack.argtypes = [C.c_void_p] * 6 + [C.c_longlong] + [C.c_void_p] * 4
    L.t_viewer_pack_kernel_emulation.argtypes = [C.c_void_p] * 6 + [C.c_longlong] + [C.c_void_p] * 4 + [C.c_int]
    L.t_f32_to_f16_glm.argtypes, L.t_f32_to_f16_glm.restype = [C.c_float], C.c_uint32
    L.t_f32_to_ordered.argtypes, L.t_f32_to_ordered.restype = [C.c_float], C.c_uint32
    L.t_ordered_to_f32.argtypes, L.t_ordered_to_f32.restype = [C.c_uint32], C.c_float
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_with_host_ops(L, m):
    N = m["opac"].shape[0]
    g, c, sh, bb = np.zeros((N, 8), np.uint32), np.zeros((N, 2), np.uint32), np.zeros((N, 16), np.uint32), np.zeros(6, np.uint32)
    L.t_viewer_pack(*[_p(np.ascontiguousarray(m[k])) for k in KEYS], N, _p(g), _p(c), _p(sh), _p(bb))
    box = np.array([L.t_ordered_to_f32(int(v)) for v in bb], np.float32)
    return g, c, sh, box


def pack_with_reference(m):
    R = C.CDLL(REF_SO)
    R.ref_viewer_pack.argtypes = [C.c_void_p] * 6 + [C.c_longlong] + [C.c_void_p] * 4
    N = m["opac"].shape[0]
    g, c, sh, box = np.zeros((N, 8), np.uint32), np.zeros((N, 2), np.uint32), np.zeros((N, 16), np.uint32), np.zeros(6, np.float32)
    assert R.ref_viewer_pack(*[_p(np.ascontiguousarray(m[k])) for k in KEYS], N, _p(g), _p(c), _p(sh), _p(box)) == 0
    return g, c, sh, box


def digest(g, c, sh, box):
    h = hashlib.sha256()
    for a in (g, c, sh, box):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()
