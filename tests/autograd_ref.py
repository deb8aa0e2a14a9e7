"""Independent float64 torch-autograd re-expression of the rasterize path (test helper).

Purpose (SURVEY.md §7 "Hard parts", §8-C G7): the reference ships no implementation, test or golden
vector for this path, so the CPU oracle's *analytic* backward could be self-consistent-but-wrong.
This file re-derives every gradient by automatic differentiation of a forward written directly from
the maths of SURVEY.md Appendix B (B.1, B.3), sharing no code with oracle/ or the CUDA kernels.  Only the
non-differentiable structure (sorted per-tile lists) is taken from the caller.

The three places where the credited algorithm's backward deliberately differs from the true derivative
are mirrored explicitly (Appendix B.4/B.5): the 0.99 alpha clamp is straight-through, the clamped
t~ of the EWA Jacobian is a constant when the clamp is active, and skip / termination decisions carry
no gradient.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def _mat(flat):
    """flat[4c+r] -> row-major 4x4 torch float64."""
    return torch.tensor(np.asarray(flat, np.float64).reshape(4, 4).T.copy())


def sh_basis(deg, d):
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    b = [torch.full_like(x, SH_C0)]
    if deg >= 1:
        b += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if deg >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
        if deg >= 3:
            b += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
                  SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy),
                  SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3 * yy)]
    return torch.stack(b, 1)  # [N,K]


def project(cam, means, log_scales, quats, logit, sh0, shN, sh_degree, activated=False, antialias=False):
    """Per-Gaussian forward (Appendix B.1) in float64; returns dict of differentiable tensors."""
    V = _mat(cam.view); PV = _mat(cam.proj)
    W, H = cam.width, cam.height
    N = means.shape[0]
    ones = torch.ones(N, 1, dtype=torch.float64)
    ph = torch.cat([means, ones], 1)
    t = (ph @ V.T)[:, :3]
    h = ph @ PV.T
    w_inv = 1.0 / (h[:, 3] + 1e-7)
    ndc = h[:, :2] * w_inv[:, None]
    if activated:
        s = log_scales * cam.scale_modifier; q = quats; o = logit
    else:
        s = torch.exp(log_scales) * cam.scale_modifier
        q = quats / quats.norm(dim=1, keepdim=True)
        o = torch.sigmoid(logit)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], 1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], 1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)], 1)
    M = R * s[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    fx = W / (2.0 * cam.tanfovx); fy = H / (2.0 * cam.tanfovy)
    limx = 1.3 * cam.tanfovx; limy = 1.3 * cam.tanfovy
    tz = t[:, 2]
    rx, ry = t[:, 0] / tz, t[:, 1] / tz
    inx = (rx >= -limx) & (rx <= limx); iny = (ry >= -limy) & (ry <= limy)
    tx = torch.where(inx, t[:, 0], (rx.clamp(-limx, limx) * tz).detach())
    ty = torch.where(iny, t[:, 1], (ry.clamp(-limy, limy) * tz).detach())
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / tz, zero, -fx * tx / (tz * tz)], 1),
                     torch.stack([zero, fy / tz, -fy * ty / (tz * tz)], 1)], 1)  # [N,2,3]
    Tm = J @ V[:3, :3]
    cov = Tm @ Sigma @ Tm.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3; b = cov[:, 0, 1]; c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    if antialias:  # mip-splatting opacity compensation (gsplat_vs.hlsl:296-301)
        det0 = cov[:, 0, 0] * cov[:, 1, 1] - b * b
        o = o * torch.sqrt(torch.clamp(det0 / det, min=0.0))
    conic = torch.stack([c / det, -b / det, a / det], 1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam))
    mean2D = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    campos = torch.tensor(np.asarray(cam.campos, np.float64))
    d = means - campos
    d = d / d.norm(dim=1, keepdim=True)
    K = (sh_degree + 1) ** 2
    bas = sh_basis(sh_degree, d)
    coeffs = torch.cat([sh0[:, None, :], shN[:, :K - 1, :]], 1) if K > 1 else sh0[:, None, :]
    col = (bas[:, :, None] * coeffs).sum(1) + 0.5
    rgb = torch.clamp(col, min=0.0)
    # F4 normal: the shortest axis of the ellipsoid, in view space, turned towards the camera (piecewise-constant choices)
    with torch.no_grad():
        jmin = s.argmin(1)
    n_v = R[torch.arange(N), :, jmin] @ V[:3, :3].T
    with torch.no_grad():
        sign = torch.where((n_v * t).sum(1) > 0, -torch.ones(N, dtype=torch.float64), torch.ones(N, dtype=torch.float64))
    return dict(t=t, depth=tz, mean2D=mean2D, conic=conic, opacity=o, rgb=rgb, radius=radius, det=det,
                cov2D=torch.stack([a, b, c], 1), Sigma=Sigma, normal=n_v * sign[:, None])


def composite(cam, proj, ranges, point_list, visible, aux=None):
    """Tile compositing (Appendix B.3) in float64 using the given sorted lists.  `aux`: a [2,H,W] or [5,H,W] tensor that
    receives the auxiliary maps depth = sum_k w_k z_k, alpha = sum_k w_k and (5 channels) normal = sum_k w_k n_k (row F4)."""
    W, H = cam.width, cam.height
    gx = (W + 15) // 16
    bg = torch.tensor(np.asarray(cam.bg, np.float64))
    img = torch.zeros(3, H, W, dtype=torch.float64) + bg[:, None, None]
    n_contrib = np.zeros((H, W), np.int64)
    final_T = np.ones((H, W), np.float64)
    pl = torch.as_tensor(np.asarray(point_list, np.int64))
    for tile in range(ranges.shape[0]):
        r0, r1 = int(ranges[tile, 0]), int(ranges[tile, 1])
        x0, y0 = (tile % gx) * 16, (tile // gx) * 16
        xs = torch.arange(x0, min(x0 + 16, W), dtype=torch.float64)
        ys = torch.arange(y0, min(y0 + 16, H), dtype=torch.float64)
        if r1 <= r0 or len(xs) == 0 or len(ys) == 0:
            continue
        ids = pl[r0:r1]
        px = xs[None, :].expand(len(ys), len(xs)).reshape(-1)
        py = ys[:, None].expand(len(ys), len(xs)).reshape(-1)
        m = proj["mean2D"][ids]
        dx = m[None, :, 0] - px[:, None]; dy = m[None, :, 1] - py[:, None]
        con = proj["conic"][ids]
        power = -0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) - con[None, :, 1] * dx * dy
        a_raw = proj["opacity"][ids][None, :] * torch.exp(power)
        alpha = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()  # straight-through 0.99 clamp
        with torch.no_grad():
            keep = (power <= 0) & (alpha >= 1.0 / 255.0) & torch.as_tensor(visible[ids.numpy()])[None, :]
            a_eff = torch.where(keep, alpha, torch.zeros_like(alpha))
            T_after = torch.cumprod(1.0 - a_eff, 1)
            stop = keep & (T_after < 1e-4)
            # first stopping index per pixel; entries at/after it do not contribute
            any_stop = stop.any(1)
            first = torch.where(any_stop, stop.float().argmax(1), torch.full((stop.shape[0],), stop.shape[1]))
            idx = torch.arange(stop.shape[1])[None, :]
            keep = keep & (idx < first[:, None])
        a_k = torch.where(keep, alpha, torch.zeros_like(alpha))
        T_excl = torch.cumprod(torch.cat([torch.ones(a_k.shape[0], 1, dtype=torch.float64), 1.0 - a_k], 1), 1)
        wgt = a_k * T_excl[:, :-1]
        col = wgt @ proj["rgb"][ids]                  # [P,3]
        T_fin = T_excl[:, -1]
        out = col + T_fin[:, None] * bg[None, :]
        ny, nx = len(ys), len(xs)
        img[:, y0:y0 + ny, x0:x0 + nx] = out.T.reshape(3, ny, nx)
        if aux is not None:
            feats = torch.stack([proj["depth"][ids], torch.ones(len(ids), dtype=torch.float64)], 1)
            if aux.shape[0] == 5:
                feats = torch.cat([feats, proj["normal"][ids]], 1)
            extra = wgt @ feats  # [P, 2 or 5]
            aux[:, y0:y0 + ny, x0:x0 + nx] = extra.T.reshape(aux.shape[0], ny, nx)
        with torch.no_grad():
            kk = keep.numpy()
            last = np.where(kk.any(1), kk.shape[1] - np.argmax(kk[:, ::-1], 1), 0)
            n_contrib[y0:y0 + ny, x0:x0 + nx] = last.reshape(ny, nx)
            final_T[y0:y0 + ny, x0:x0 + nx] = T_fin.numpy().reshape(ny, nx)
    return img, n_contrib, final_T


def render_and_grad(cam, scene_arrays, sh_degree, ranges, point_list, radii, dL_dpix, activated=False, antialias=False,
                    dL_daux=None):
    """Returns (image float64 numpy, dict of gradient numpy arrays w.r.t. the stored parameters).  With dL_daux [2,H,W]
    (or [5,H,W]) the loss also has <depth, dL_daux[0]> + <alpha, dL_daux[1]> (+ <normal, dL_daux[2:5]>); the maps are
    returned under proj["aux"]."""
    names = ["means3D", "scales", "quats", "opac", "sh0", "shN"]
    ts = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in zip(names, scene_arrays)}
    proj = project(cam, ts["means3D"], ts["scales"], ts["quats"], ts["opac"].reshape(-1), ts["sh0"],
                   ts["shN"], sh_degree, activated, antialias)
    visible = np.asarray(radii) > 0
    aux = torch.zeros(np.asarray(dL_daux).shape[0], cam.height, cam.width, dtype=torch.float64) if dL_daux is not None else None
    img, n_contrib, final_T = composite(cam, proj, np.asarray(ranges), point_list, visible, aux)
    loss = (img * torch.tensor(np.asarray(dL_dpix, np.float64))).sum()
    if aux is not None:
        loss = loss + (aux * torch.tensor(np.asarray(dL_daux, np.float64))).sum()
        proj["aux"] = aux.detach().numpy()
    loss.backward()
    grads = {k: (ts[k].grad.numpy() if ts[k].grad is not None else np.zeros_like(ts[k].detach().numpy()))
             for k in names}
    return img.detach().numpy(), grads, proj, n_contrib, final_T
