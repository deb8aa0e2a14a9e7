"""Host-side front-end of the rasterizer used by tests and bench.py (torch tensors in, C-ABI underneath).

`Rasterizer` mirrors the operator surface of the absent `diverse_utils/gsplatrast` libtorch operator
(SURVEY.md §8 A9): forward(camera, params) -> (image [3,H,W], radii [N]); backward(dL_dpix) -> gradients
for means3D / scales / rotations / opacity / sh0 / shN (+ the screen-space statistics the trainer's densify
uses).  torch is only plumbing here: device memory, streams and autograd bookkeeping.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi

PARAM_NAMES = ("means3D", "scales", "quats", "opacities", "sh0", "shN")


class RasterizerError(RuntimeError):
    pass


@dataclass
class GradBuffers:
    """One contiguous fp32 arena [N*(11+3K')] with six views — a single NCCL all-reduce covers it."""
    flat: torch.Tensor
    means3D: torch.Tensor
    scales: torch.Tensor
    quats: torch.Tensor
    opacities: torch.Tensor
    sh0: torch.Tensor
    shN: torch.Tensor

    @staticmethod
    def numel_for(N: int, sh_rest: int) -> int:
        total = 0
        for sz in (4 * N, 3 * sh_rest * N, 3 * N, 3 * N, 3 * N, N):
            total = (total + 3) // 4 * 4 + sz
        return max(total, 4)

    @staticmethod
    def allocate(N: int, sh_rest: int, device, flat: torch.Tensor | None = None) -> "GradBuffers":
        # every view starts on a 16-byte boundary: order quats, shN first (16 B rows), then the 12 B / 4 B rows
        sizes = [("quats", 4 * N), ("shN", 3 * sh_rest * N), ("means3D", 3 * N), ("scales", 3 * N), ("sh0", 3 * N),
                 ("opacities", N)]
        offs, total = {}, 0
        for name, sz in sizes:
            total = (total + 3) // 4 * 4
            offs[name] = (total, sz)
            total += sz
        if flat is None:
            flat = torch.zeros(max(total, 4), dtype=torch.float32, device=device)
        assert flat.numel() >= total and flat.is_contiguous()
        v = {n: flat[o:o + s] for n, (o, s) in offs.items()}
        return GradBuffers(flat, v["means3D"].view(N, 3), v["scales"].view(N, 3), v["quats"].view(N, 4),
                           v["opacities"].view(N), v["sh0"].view(N, 3), v["shN"].view(N, sh_rest, 3))


class Rasterizer:
    def __init__(self, device: int | torch.device = 0):
        if not torch.cuda.is_available():
            raise RasterizerError("CUDA is not available: the rasterizer has no CPU path")
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        self._lib = _cabi.load()
        h = C.c_void_p()
        rc = self._lib.dvs_rast_create(self.device.index, C.byref(h))
        if rc != 0:
            raise RasterizerError(f"dvs_rast_create failed ({rc})")
        self._h = h
        self._cam = None
        self._params = None
        # the forward state lives in the context, one copy: autograd nodes remember the forward they belong to
        self.generation = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dvs_rast_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise RasterizerError(f"dvs_rast error {rc}: {self._lib.dvs_rast_last_error(self._h).decode()}")

    @staticmethod
    def _pstruct(params) -> _cabi.DvsParams:
        p = _cabi.DvsParams()
        for n in PARAM_NAMES:
            t = params[n]
            if t is not None and t.numel() > 0:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), n
                setattr(p, n, t.data_ptr())
        return p

    def reserve(self, N, width, height, dup_capacity=0):
        self._check(self._lib.dvs_rast_reserve(self._h, N, width, height, dup_capacity))

    def forward(self, cam: _cabi.DvsCamera, params: dict, out_color: torch.Tensor | None = None,
                out_radii: torch.Tensor | None = None, defer_check: bool = False):
        """defer_check: no host synchronisation in the call (DVS_FLAG_DEFER_CHECK); a binning-arena overflow is then
        raised by a later call (RasterizerError, code DVS_E_OVERFLOW) and the step must be redone."""
        N = params["means3D"].shape[0]
        if defer_check != bool(cam.flags & _cabi.FLAG_DEFER_CHECK):
            cam2 = _cabi.DvsCamera.from_buffer_copy(cam)
            cam2.flags = (cam.flags | _cabi.FLAG_DEFER_CHECK) if defer_check else (cam.flags & ~_cabi.FLAG_DEFER_CHECK)
            cam = cam2
        if out_color is None:
            out_color = torch.empty(3, cam.height, cam.width, dtype=torch.float32, device=self.device)
        if out_radii is None:
            out_radii = torch.empty(N, dtype=torch.int32, device=self.device)
        self._cam, self._params = cam, params
        self.generation += 1
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.dvs_rast_forward(self._h, C.byref(cam), N, C.byref(self._pstruct(params)),
                                               out_color.data_ptr(), out_radii.data_ptr(), C.c_void_p(st)))
        return out_color, out_radii

    def backward(self, dL_dpix: torch.Tensor, grads: GradBuffers, flags: int = 0,
                 mean2D: torch.Tensor | None = None, mean2D_abs: torch.Tensor | None = None):
        assert dL_dpix.is_cuda and dL_dpix.dtype == torch.float32 and dL_dpix.is_contiguous()
        g = _cabi.DvsGrads()
        for n in PARAM_NAMES:
            t = getattr(grads, n)
            if t.numel() > 0:
                setattr(g, n, t.data_ptr())
        if mean2D is not None:
            g.mean2D = mean2D.data_ptr()
        if mean2D_abs is not None:
            g.mean2D_abs = mean2D_abs.data_ptr()
            flags |= _cabi.FLAG_ABSGRAD
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.dvs_rast_backward(self._h, C.byref(self._pstruct(self._params)), dL_dpix.data_ptr(),
                                                C.byref(g), flags, C.c_void_p(st)))
        return grads

    def forward_aux(self, normals: bool = False):
        """(depth, alpha) [2,H,W] of the last forward, and with normals=True also the normal map [3,H,W]
        (dvs_rast_forward_aux)."""
        H, W = self._cam.height, self._cam.width
        out_aux = torch.empty(2, H, W, dtype=torch.float32, device=self.device)
        out_n = torch.empty(3, H, W, dtype=torch.float32, device=self.device) if normals else None
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.dvs_rast_forward_aux(self._h, C.byref(self._pstruct(self._params)), out_aux.data_ptr(),
                                                   out_n.data_ptr() if normals else None, C.c_void_p(st)))
        return (out_aux, out_n) if normals else out_aux

    def backward_aux(self, dL_dpix: torch.Tensor, dL_daux: torch.Tensor | None, grads: GradBuffers, flags: int = 0,
                     dL_dnormal: torch.Tensor | None = None):
        """Backward of <image, dL_dpix> + <depth, dL_daux[0]> + <alpha, dL_daux[1]> + <normal map, dL_dnormal>
        (dvs_rast_backward_aux)."""
        for t in (dL_dpix, dL_daux, dL_dnormal):
            assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
        g = _cabi.DvsGrads()
        for n in PARAM_NAMES:
            t = getattr(grads, n)
            if t.numel() > 0:
                setattr(g, n, t.data_ptr())
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.dvs_rast_backward_aux(self._h, C.byref(self._pstruct(self._params)), dL_dpix.data_ptr(),
                                                    dL_daux.data_ptr() if dL_daux is not None else None,
                                                    dL_dnormal.data_ptr() if dL_dnormal is not None else None,
                                                    C.byref(g), flags, C.c_void_p(st)))
        return grads

    def step_host(self, cam, params: dict, grads: GradBuffers, dL_dpix_host: torch.Tensor,
                  out_color_host: torch.Tensor, flags: int = 0):
        """End-to-end step with HOST image buffers (pinned): H2D dL/dpix, forward, backward, D2H image."""
        N = params["means3D"].shape[0]
        g = _cabi.DvsGrads()
        for n in PARAM_NAMES:
            t = getattr(grads, n)
            if t.numel() > 0:
                setattr(g, n, t.data_ptr())
        self._cam, self._params = cam, params
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.dvs_rast_step_host(self._h, C.byref(cam), N, C.byref(self._pstruct(params)), C.byref(g),
                                                 dL_dpix_host.data_ptr(), out_color_host.data_ptr(), flags,
                                                 C.c_void_p(st)))

    def step_host_async(self, cam, params: dict, grads: GradBuffers, dL_dpix_host: torch.Tensor,
                        out_color_host: torch.Tensor, slot: int, flags: int = 0):
        """Pipelined end-to-end step (dvs_rast_step_host_async): queues H2D, forward, D2H, backward on `slot` and returns;
        step_host_wait(slot) blocks until that step's image is in out_color_host."""
        N = params["means3D"].shape[0]
        g = _cabi.DvsGrads()
        for n in PARAM_NAMES:
            t = getattr(grads, n)
            if t.numel() > 0:
                setattr(g, n, t.data_ptr())
        self._cam, self._params = cam, params
        self.generation += 1
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.dvs_rast_step_host_async(self._h, C.byref(cam), N, C.byref(self._pstruct(params)), C.byref(g),
                                                       dL_dpix_host.data_ptr(), out_color_host.data_ptr(), flags, slot,
                                                       C.c_void_p(st)))

    def step_host_wait(self, slot: int):
        self._check(self._lib.dvs_rast_step_host_wait(self._h, slot))

    def kernel_launches(self) -> int:
        """Kernels of the library enqueued through this context so far (dvs_rast_kernel_launches)."""
        return int(self._lib.dvs_rast_kernel_launches(self._h))

    def set_background(self, bg_image: torch.Tensor | None):
        """Per-pixel background [3,H,W] (the trainer's sky model, GaussianTrainConfig::enableBg) for every later forward /
        backward of this context; None restores the camera's constant background.  The tensor is kept alive here."""
        if bg_image is not None:
            assert bg_image.is_cuda and bg_image.dtype == torch.float32 and bg_image.is_contiguous() and bg_image.dim() == 3
        self._bg = bg_image
        self._check(self._lib.dvs_rast_set_background(self._h, bg_image.data_ptr() if bg_image is not None else None))

    def background_grad(self, dL_dpix: torch.Tensor) -> torch.Tensor:
        """dL/dbg = final_T * dL/dpix [3,H,W] of the last forward."""
        out = torch.empty_like(dL_dpix)
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.dvs_rast_background_grad(self._h, dL_dpix.data_ptr(), out.data_ptr(), C.c_void_p(st)))
        return out

    def set_profiling(self, on: bool):
        """Per-stage CUDA events on/off (off in a training loop; stage_ms() needs them on)."""
        if hasattr(self._lib, "dvs_rast_set_profiling"):
            self._check(self._lib.dvs_rast_set_profiling(self._h, int(on)))

    def stats(self) -> dict:
        s = _cabi.DvsStats()
        self._check(self._lib.dvs_rast_get_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in s._fields_}

    def stage_ms(self) -> dict:
        arr = (C.c_float * _cabi.NUM_STAGES)()
        self._check(self._lib.dvs_rast_stage_ms(self._h, C.byref(arr)))
        return {self._lib.dvs_rast_stage_name(i).decode(): float(arr[i]) for i in range(_cabi.NUM_STAGES)}

    def debug_read(self, which: int) -> np.ndarray:
        s = self.stats()
        N, D = s["num_gaussians"], s["num_list_entries"]
        T = s["tiles_x"] * s["tiles_y"]
        P = self._cam.width * self._cam.height
        shapes = {
            _cabi.BUF_RADII: ((N,), np.int32), _cabi.BUF_TILES_TOUCHED: ((N,), np.uint32),
            _cabi.BUF_DEPTH: ((N,), np.float32), _cabi.BUF_MEAN2D: ((N, 2), np.float32),
            _cabi.BUF_CONIC_OPACITY: ((N, 4), np.float32), _cabi.BUF_RGB: ((N, 3), np.float32),
            _cabi.BUF_CLAMPED: ((N, 3), np.uint8), _cabi.BUF_POINT_LIST: ((D,), np.uint32),
            _cabi.BUF_RANGES: ((T, 2), np.uint32), _cabi.BUF_FINAL_T: ((P,), np.float32),
            _cabi.BUF_N_CONTRIB: ((P,), np.uint32), _cabi.BUF_CULL_MASK: ((D,), np.uint8),
            _cabi.BUF_SCREEN_GRADS: ((N, 12), np.float32),
        }
        shape, dt = shapes[which]
        out = np.zeros(shape, dt)
        self._check(self._lib.dvs_rast_debug_read(self._h, which, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out


def scene_to_device(scene, device) -> dict:
    """Upload a scenes.Scene to CUDA tensors (the trainer's resident parameters)."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return {"means3D": t(scene.means3D), "scales": t(scene.log_scales), "quats": t(scene.quats),
            "opacities": t(scene.logit_opac), "sh0": t(scene.sh0),
            "shN": t(scene.shN) if scene.shN.size else torch.zeros(scene.N, 0, 3, device=device)}


class _RasterizeFn(torch.autograd.Function):
    """autograd bridge: gradients flow to the six parameter tensors."""

    @staticmethod
    def forward(ctx, rast, cam, means3D, scales, quats, opacities, sh0, shN):
        params = dict(means3D=means3D.contiguous(), scales=scales.contiguous(), quats=quats.contiguous(),
                      opacities=opacities.contiguous().view(-1), sh0=sh0.contiguous(), shN=shN.contiguous())
        img, radii = rast.forward(cam, params)
        ctx.rast, ctx.params, ctx.shapes = rast, params, (opacities.shape, shN.shape)
        ctx.generation = rast.generation
        ctx.mark_non_differentiable(radii)
        return img, radii

    @staticmethod
    def backward(ctx, dL_dimg, _):
        if ctx.generation != ctx.rast.generation:
            raise RasterizerError(
                f"backward of forward #{ctx.generation} but the context holds the state of forward #{ctx.rast.generation}: a "
                "Rasterizer keeps ONE outstanding forward (tile lists, final_T, camera); render concurrent views with "
                "separate Rasterizers")
        N = ctx.params["means3D"].shape[0]
        g = GradBuffers.allocate(N, ctx.shapes[1][1], dL_dimg.device)
        ctx.rast._params = ctx.params
        ctx.rast.backward(dL_dimg.contiguous(), g)
        return (None, None, g.means3D, g.scales, g.quats, g.opacities.view(ctx.shapes[0]), g.sh0, g.shN)


def rasterize_gaussians(rast: Rasterizer, cam, means3D, scales, quats, opacities, sh0, shN):
    return _RasterizeFn.apply(rast, cam, means3D, scales, quats, opacities, sh0, shN)
