"""Synthetic scenes for the BASELINE.json configs (SURVEY.md §8-D "Synthetic scene").

Deterministic: seed = 1234 + config index, torch.Generator on CPU, fp32.  No dataset exists in the
reference, so every parity test and bench line runs on these.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

# name: (index, N, W, H, sh_degree, views)
CONFIGS = {
    "c1": (1, 10_000, 256, 256, 0, 1),
    "c2": (2, 100_000, 800, 600, 1, 1),
    "c3": (3, 1_000_000, 1600, 1000, 3, 1),
    "c4": (4, 1_000_000, 1600, 1000, 3, 8),
    "c5": (5, 5_000_000, 1920, 1080, 3, 1),
}


@dataclass
class Camera:
    view: np.ndarray      # flat float32[16], [4*c + r] = row r, col c (world -> view)
    proj: np.ndarray      # flat float32[16], full view-projection, same layout
    campos: np.ndarray    # float32[3]
    tanfovx: float
    tanfovy: float
    width: int
    height: int
    bg: np.ndarray        # float32[3]
    scale_modifier: float = 1.0


@dataclass
class Scene:
    name: str
    means3D: np.ndarray   # [N,3]
    log_scales: np.ndarray  # [N,3] raw (exp activation)
    quats: np.ndarray     # [N,4] r,x,y,z raw
    logit_opac: np.ndarray  # [N] raw (sigmoid activation)
    sh0: np.ndarray       # [N,3]
    shN: np.ndarray       # [N,K-1,3]
    sh_degree: int
    cameras: list
    dL_dpix: list         # per view [3,H,W]

    @property
    def N(self):
        return self.means3D.shape[0]

    @property
    def K(self):
        return (self.sh_degree + 1) ** 2

    def param_bytes(self):
        return 44 + 12 * self.K


def _flat(M: np.ndarray) -> np.ndarray:
    """Row-major 4x4 -> flat layout with element [4*c + r] = M[r, c]."""
    return np.ascontiguousarray(M.T.reshape(16).astype(np.float32))


def look_at_camera(eye, target, width, height, fovx_deg=60.0, znear=0.01, zfar=100.0, bg=(0, 0, 0)) -> Camera:
    """+z-forward, y-down COLMAP-style camera (the trainer's convention, editor.cpp:2028-2029)."""
    eye = np.asarray(eye, np.float64); target = np.asarray(target, np.float64)
    f = target - eye; f /= np.linalg.norm(f)
    up = np.array([0.0, 1.0, 0.0])
    r = np.cross(up, f)
    if np.linalg.norm(r) < 1e-8:
        r = np.array([1.0, 0.0, 0.0])
    r /= np.linalg.norm(r)
    u = np.cross(f, r)
    R = np.stack([r, u, f], 0)  # world -> view rotation rows
    V = np.eye(4); V[:3, :3] = R; V[:3, 3] = -R @ eye
    tanfovx = math.tan(math.radians(fovx_deg) * 0.5)
    tanfovy = tanfovx * height / width
    P = np.zeros((4, 4))
    P[0, 0] = 1.0 / tanfovx
    P[1, 1] = 1.0 / tanfovy
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    P[3, 2] = 1.0
    PV = P @ V
    return Camera(_flat(V), _flat(PV), eye.astype(np.float32), float(np.float32(tanfovx)),
                  float(np.float32(tanfovy)), int(width), int(height), np.asarray(bg, np.float32))


def make_scene(name: str = None, *, N=None, width=None, height=None, sh_degree=None, views=None,
               seed=None, normalise_quats=True, bg=(0, 0, 0), with_grad=True, tanfovx=None, mu_s=None) -> Scene:
    """tanfovx / mu_s override the 60-degree FOV and the N-dependent mean log-scale (used for
    density-preserving crops of a config: same focal length and splat size, fewer splats and pixels)."""
    if name is not None and name in CONFIGS:
        idx, N0, W0, H0, d0, v0 = CONFIGS[name]
        N = N0 if N is None else N; width = W0 if width is None else width
        height = H0 if height is None else height
        sh_degree = d0 if sh_degree is None else sh_degree; views = v0 if views is None else views
        seed = 1234 + idx if seed is None else seed
    else:
        views = 1 if views is None else views
        seed = 1234 if seed is None else seed
        name = name or f"custom_{N}_{width}x{height}_d{sh_degree}"
    g = torch.Generator(device="cpu"); g.manual_seed(int(seed))
    tanfovx = math.tan(math.radians(60.0) * 0.5) if tanfovx is None else float(tanfovx)
    fovx_deg = math.degrees(2.0 * math.atan(tanfovx))
    tanfovy = tanfovx * height / width
    z = torch.empty(N).uniform_(2.0, 10.0, generator=g)
    ux = torch.empty(N).uniform_(-1.15, 1.15, generator=g)
    uy = torch.empty(N).uniform_(-1.15, 1.15, generator=g)
    means = torch.stack([ux * tanfovx * z, uy * tanfovy * z, z], 1)
    mu_s = math.log(0.012 * (1.0e6 / N) ** (1.0 / 3.0)) if mu_s is None else float(mu_s)
    log_scales = torch.randn(N, 3, generator=g) * 0.5 + mu_s
    quats = torch.randn(N, 4, generator=g)
    if normalise_quats:
        quats = quats / quats.norm(dim=1, keepdim=True)
    logit = torch.randn(N, generator=g) * 1.5
    K = (sh_degree + 1) ** 2
    sh0 = torch.randn(N, 3, generator=g)
    shN = torch.randn(N, K - 1, 3, generator=g) * 0.2
    cams, grads = [], []
    for v in range(views):
        if views == 1:
            eye, tgt = (0.0, 0.0, 0.0), (0.0, 0.0, 6.0)
        else:
            ang = 2.0 * math.pi * v / views
            eye, tgt = (0.5 * math.cos(ang), 0.5 * math.sin(ang), 0.0), (0.0, 0.0, 6.0)
        cams.append(look_at_camera(eye, tgt, width, height, fovx_deg=fovx_deg, bg=bg))
        if with_grad:
            gv = torch.Generator(device="cpu"); gv.manual_seed(int(seed) * 1000 + v)
            grads.append(torch.randn(3, height, width, generator=gv).numpy())
    f = lambda t: np.ascontiguousarray(t.numpy().astype(np.float32))
    return Scene(name, f(means), f(log_scales), f(quats), f(logit), f(sh0), f(shN), int(sh_degree), cams, grads)


def crop_of(name: str, frac_lin: int, seed=None) -> Scene:
    """Density-preserving crop of a config: image and FOV shrunk by `frac_lin` per axis, N by frac_lin^2,
    same focal length and the config's own splat-size distribution.  Used as the bounded CPU sample."""
    idx, N0, W0, H0, d0, _ = CONFIGS[name]
    tan0 = math.tan(math.radians(60.0) * 0.5)
    return make_scene(None, N=N0 // (frac_lin * frac_lin), width=W0 // frac_lin, height=H0 // frac_lin,
                      sh_degree=d0, seed=(1234 + idx if seed is None else seed), tanfovx=tan0 / frac_lin,
                      mu_s=math.log(0.012 * (1.0e6 / N0) ** (1.0 / 3.0)))
