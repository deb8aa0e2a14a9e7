"""Independent float64 torch-autograd re-expression of the 2DGS ("surfel", GaussianTrainConfig::modelType = 1) rasterize
path — test helper, shares no code with oracle/ or the CUDA kernels.  Written from the published algorithm (the S.1 / S.2
comment of oracle/dvs_oracle.c restates it): a Gaussian is a flat disk with tangents R[:,0], R[:,1] and scales (s_u, s_v); its
homography M = Npix Proj [s_u t_u | s_v t_v | p; 0 0 1] maps local (u, v, 1) to homogeneous pixel coordinates; per pixel the
ray-splat intersection (u, v) = ((x Tw - Tu) x (y Tw - Tv)).xy / .z gives rho3d = u^2 + v^2, the object-space low-pass filter
rho2d = 2 |c - pixel|^2 with c the projected centre computed FROM M, alpha = min(0.99, o exp(-min(rho3d, rho2d) / 2)), then
the usual front-to-back compositing.  Only the non-differentiable structure (sorted per-tile lists) comes from the caller."""
from __future__ import annotations

import numpy as np
import torch

from autograd_ref import _mat, sh_basis


def project2d(cam, means, log_scales, quats, logit, sh0, shN, sh_degree):
    V = _mat(cam.view); PV = _mat(cam.proj)
    W, H = cam.width, cam.height
    N = means.shape[0]
    ph = torch.cat([means, torch.ones(N, 1, dtype=torch.float64)], 1)
    t = (ph @ V.T)[:, :3]
    s = torch.exp(log_scales) * cam.scale_modifier
    q = quats / quats.norm(dim=1, keepdim=True)
    o = torch.sigmoid(logit)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], 1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], 1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)], 1)
    zero = torch.zeros(N, 1, dtype=torch.float64)
    A = torch.stack([torch.cat([R[:, :, 0] * s[:, 0:1], zero], 1), torch.cat([R[:, :, 1] * s[:, 1:2], zero], 1), ph], 2)  # [N,4,3]
    Npix = torch.tensor([[W / 2.0, 0, 0, (W - 1) / 2.0], [0, H / 2.0, 0, (H - 1) / 2.0], [0, 0, 0, 1.0]], dtype=torch.float64)
    M = (Npix @ PV)[None] @ A                                  # [N,3,3]: rows Tu, Tv, Tw
    Tu, Tv, Tw = M[:, 0], M[:, 1], M[:, 2]
    tp = torch.tensor([9.0, 9.0, -1.0], dtype=torch.float64)
    dist = (tp * Tw * Tw).sum(1)
    f = tp[None] / dist[:, None]
    c = torch.stack([(f * Tu * Tw).sum(1), (f * Tv * Tw).sum(1)], 1)
    campos = torch.tensor(np.asarray(cam.campos, np.float64))
    d = means - campos
    d = d / d.norm(dim=1, keepdim=True)
    K = (sh_degree + 1) ** 2
    bas = sh_basis(sh_degree, d)
    coeffs = torch.cat([sh0[:, None, :], shN[:, :K - 1, :]], 1) if K > 1 else sh0[:, None, :]
    rgb = torch.clamp((bas[:, :, None] * coeffs).sum(1) + 0.5, min=0.0)
    return dict(Tu=Tu, Tv=Tv, Tw=Tw, center=c, opacity=o, rgb=rgb, depth=t[:, 2])


def composite2d(cam, proj, ranges, point_list, visible):
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
        px = xs[None, :].expand(len(ys), len(xs)).reshape(-1)[:, None, None]
        py = ys[:, None].expand(len(ys), len(xs)).reshape(-1)[:, None, None]
        Tu, Tv, Tw = proj["Tu"][ids][None], proj["Tv"][ids][None], proj["Tw"][ids][None]   # [1,n,3]
        k = px * Tw - Tu; l = py * Tw - Tv                                                  # [P,n,3]
        pv = torch.cross(k, l, dim=2)
        pz = pv[..., 2]
        safe = torch.where(pz == 0, torch.ones_like(pz), pz)
        u, v = pv[..., 0] / safe, pv[..., 1] / safe
        rho3d = u * u + v * v
        c = proj["center"][ids][None]
        dx, dy = c[..., 0] - px[..., 0], c[..., 1] - py[..., 0]
        rho2d = 2.0 * (dx * dx + dy * dy)
        rho = torch.minimum(rho3d, rho2d)
        with torch.no_grad():
            use3d = rho3d <= rho2d
            dep = torch.where(use3d, u * Tw[..., 0] + v * Tw[..., 1] + Tw[..., 2], Tw[..., 2].expand_as(u))
        a_raw = proj["opacity"][ids][None, :] * torch.exp(-0.5 * rho)
        alpha = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()
        with torch.no_grad():
            keep = (pz != 0) & (dep >= 0.2) & (alpha >= 1.0 / 255.0) & torch.as_tensor(visible[ids.numpy()])[None, :]
            a_eff = torch.where(keep, alpha, torch.zeros_like(alpha))
            T_after = torch.cumprod(1.0 - a_eff, 1)
            stop = keep & (T_after < 1e-4)
            any_stop = stop.any(1)
            first = torch.where(any_stop, stop.float().argmax(1), torch.full((stop.shape[0],), stop.shape[1]))
            idx = torch.arange(stop.shape[1])[None, :]
            keep = keep & (idx < first[:, None])
        a_k = torch.where(keep, alpha, torch.zeros_like(alpha))
        T_excl = torch.cumprod(torch.cat([torch.ones(a_k.shape[0], 1, dtype=torch.float64), 1.0 - a_k], 1), 1)
        wgt = a_k * T_excl[:, :-1]
        col = wgt @ proj["rgb"][ids]
        T_fin = T_excl[:, -1]
        out = col + T_fin[:, None] * bg[None, :]
        ny, nx = len(ys), len(xs)
        img[:, y0:y0 + ny, x0:x0 + nx] = out.T.reshape(3, ny, nx)
        with torch.no_grad():
            kk = keep.numpy()
            last = np.where(kk.any(1), kk.shape[1] - np.argmax(kk[:, ::-1], 1), 0)
            n_contrib[y0:y0 + ny, x0:x0 + nx] = last.reshape(ny, nx)
            final_T[y0:y0 + ny, x0:x0 + nx] = T_fin.numpy().reshape(ny, nx)
    return img, n_contrib, final_T


def render_and_grad_2d(cam, scene_arrays, sh_degree, ranges, point_list, radii, dL_dpix):
    names = ["means3D", "scales", "quats", "opac", "sh0", "shN"]
    ts = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in zip(names, scene_arrays)}
    proj = project2d(cam, ts["means3D"], ts["scales"], ts["quats"], ts["opac"].reshape(-1), ts["sh0"], ts["shN"], sh_degree)
    visible = np.asarray(radii) > 0
    img, n_contrib, final_T = composite2d(cam, proj, np.asarray(ranges), point_list, visible)
    loss = (img * torch.tensor(np.asarray(dL_dpix, np.float64))).sum()
    loss.backward()
    grads = {k: (ts[k].grad.numpy() if ts[k].grad is not None else np.zeros_like(ts[k].detach().numpy())) for k in names}
    return img.detach().numpy(), grads, proj, n_contrib, final_T
