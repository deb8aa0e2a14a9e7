"""View-sharded data parallelism (SURVEY.md §8 e): Gaussians replicated on every rank, the views of a batch
partitioned across ranks, ONE sum all-reduce of the dense per-Gaussian gradient arena per step.

The reference has no multi-GPU path at all (zero hits for nccl / MPI / ProcessGroup); this is new
functionality required by BASELINE.json's north_star.  torch.distributed is only the plumbing (NCCL over
NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def views_for_rank(num_views: int, rank: int, world: int) -> list[int]:
    """Round-robin partition of a batch of views: rank r renders views r, r+world, ... (independent units,
    no data-path exchange until the gradient sum)."""
    return list(range(rank, num_views, world))


def allreduce_gradients(flat: torch.Tensor, group=None, average: bool = False) -> torch.Tensor:
    """Sum (or average) the flat gradient arena (GradBuffers.flat: all six parameter gradients in one
    contiguous fp32 buffer, so the exchange is a single collective call)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(dist.get_world_size(group))
    return flat


class GradientReducer:
    """Owns the flat gradient arena and the collective that sums it.

    backend "nccl": ncclAllReduce (the baseline).
    backend "nvls": the arena lives in symmetric memory (torch.distributed._symmetric_memory) and is summed by the
        NVSwitch itself with our own two-shot multimem kernel (csrc/collective.cu), bracketed by symmetric-memory
        barriers.  Each GPU moves ~(1+1/W)x the arena per direction instead of the ring's 2(W-1)/W x.
    backend "auto": NCCL unless the NVLS path initialises AND measures faster on this machine (3 timed runs each).
    """

    def __init__(self, numel: int, device: torch.device, backend: str = "auto", group=None, ctas: int = 0):
        self.group = group
        self.backend = "nccl"
        self.flat = None
        self.ctas = ctas
        self.note = ""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        numel = (numel + 1023) // 1024 * 1024
        self._hdl = None
        if world > 1 and backend in ("auto", "nvls") and device.type == "cuda":
            try:
                import ctypes as C

                import torch.distributed._symmetric_memory as symm_mem

                from . import _cabi
                self._lib = _cabi.load()
                self.group_name = (group or dist.group.WORLD).group_name
                buf = symm_mem.empty(numel, dtype=torch.float32, device=device)
                hdl = symm_mem.rendezvous(buf, self.group_name)
                if not getattr(hdl, "multicast_ptr", 0):
                    raise RuntimeError("no NVLS multicast mapping for this allocation")
                self._hdl, self._C = hdl, C
                self.flat = buf
                buf.zero_()
                self.backend = "nvls"
                self.all_reduce()
                torch.cuda.synchronize(device)
                if backend == "auto":  # keep NVLS only if it beats NCCL here
                    t_nvls = self._time(device)
                    self.backend = "nccl"
                    t_nccl = self._time(device)
                    self.backend = "nvls" if t_nvls < t_nccl else "nccl"
                    self.note = f"auto: nvls {t_nvls:.3f} ms vs nccl {t_nccl:.3f} ms"
            except Exception as e:  # no NVSwitch multicast / unsupported build: NCCL
                if backend == "nvls":
                    raise
                self.note = f"nvls unavailable ({type(e).__name__}: {e})"
                self.backend = "nccl"
        if self.flat is None:
            self.flat = torch.zeros(numel, dtype=torch.float32, device=device)

    def _time(self, device, reps: int = 3) -> float:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.all_reduce(); torch.cuda.synchronize(device)
        dist.barrier(self.group); torch.cuda.synchronize(device)
        e0.record()
        for _ in range(reps):
            self.all_reduce()
        e1.record(); torch.cuda.synchronize(device)
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def all_reduce(self):
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return self.flat
        if self.backend == "nvls":
            hdl = self._hdl
            st = torch.cuda.current_stream(self.flat.device).cuda_stream
            hdl.barrier(channel=0)  # every rank's gradients are in place
            rc = self._lib.dvs_coll_allreduce_nvls(self._C.c_void_p(hdl.multicast_ptr), self.flat.numel(), hdl.rank,
                                                   hdl.world_size, self.ctas, self._C.c_void_p(st))
            if rc != 0:
                raise RuntimeError(f"dvs_coll_allreduce_nvls failed ({rc})")
            hdl.barrier(channel=1)  # every shard has been re-broadcast
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat


class FactoredGradientExchange:
    """The gradient exchange with the SH tensor factored out (SURVEY.md §8 e; csrc/sh_grad_ops.h explains the identity).

    76 % of the dense gradient arena is dL/dshN (180 of 236 B per Gaussian at SH degree 3), yet per view it is the outer
    product B(dir) (x) dL/dcolour and the band-0 gradient already carries dL/dcolour.  So instead of all-reducing the whole
    arena, every rank
      1. all-gathers each view's dL/dsh0 (12 B per Gaussian and view) and camera centre,
      2. all-reduces everything except shN (quats | means, scales, sh0, opacities: 56 B per Gaussian, two contiguous ranges),
      3. forms sum_v B(dir_v) (x) dL/dsh0_v / SH_C0 locally (dvs_coll_sh_grad_from_dsh0) into grads.shN.
    Bytes received per GPU and Gaussian: (W-1) * 12 + 2 (W-1)/W * 56 instead of 2 (W-1)/W * 236 (8 ranks: 182 vs 413).
    Valid with ONE view per rank and step (each rank's dL/dsh0 must belong to a single camera).
    `accumulate(means, campos_all, dsh0_all, deg, out_shN)` defaults to the CUDA kernel; the CPU tests inject a stand-in.
    """

    def __init__(self, grads, group=None, accumulate=None):
        self.g, self.group = grads, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        flat = grads.flat
        N = grads.opacities.shape[0]
        self.N = N
        off = lambda t: (t.data_ptr() - flat.data_ptr()) // 4  # noqa: E731
        if grads.shN.numel() == 0:
            raise ValueError("FactoredGradientExchange: no higher SH bands (degree 0), nothing to factor out")
        # arena order: quats | shN | means3D | scales | sh0 | opacities (rasterizer.GradBuffers.allocate)
        if not (off(grads.quats) < off(grads.shN) < off(grads.means3D) < off(grads.scales) < off(grads.sh0) < off(grads.opacities)):
            raise ValueError("FactoredGradientExchange: unexpected gradient arena layout")
        self.range_a = flat[off(grads.quats):off(grads.quats) + 4 * N]
        self.range_b = flat[off(grads.means3D):off(grads.opacities) + N]
        self.dsh0_all = torch.zeros(self.world, N, 3, dtype=torch.float32, device=flat.device)
        self.campos_all = torch.zeros(self.world, 3, dtype=torch.float32, device=flat.device)
        self._accumulate = accumulate or self._accumulate_cuda

    def _accumulate_cuda(self, means, campos_all, dsh0_all, deg, out_shN):
        if not out_shN.is_cuda:
            raise RuntimeError("FactoredGradientExchange: the SH accumulation runs on the GPU only (no CPU path)")
        import ctypes as C

        from . import _cabi
        cam_host = getattr(self, "_cam_host", None)  # V x 3 floats on the host; the kernel takes them by value
        if cam_host is None:
            cam_host = campos_all.cpu().contiguous()
        st = torch.cuda.current_stream(out_shN.device).cuda_stream
        rc = _cabi.load().dvs_coll_sh_grad_from_dsh0(means.data_ptr(), cam_host.data_ptr(), dsh0_all.data_ptr(), self.N,
                                                     dsh0_all.shape[0], deg, out_shN.shape[1], out_shN.data_ptr(), C.c_void_p(st))
        if rc != 0:
            raise RuntimeError(f"dvs_coll_sh_grad_from_dsh0 failed ({rc})")

    def set_cameras(self, campos_local: torch.Tensor):
        """Camera centres change per step only if the views do; gather them once per view assignment."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.campos_all.view(-1), campos_local.to(self.campos_all.device).reshape(3).contiguous(),
                                        group=self.group)
        else:
            self.campos_all[0] = campos_local
        self._cam_host = self.campos_all.cpu().contiguous()  # cached: no device->host copy (and no stream sync) per step
        self._cam_local = campos_local.detach().to("cpu", torch.float32).reshape(3).clone()
        self._cams_set = True

    def exchange(self, means: torch.Tensor, campos_local: torch.Tensor, deg: int):
        if self.world == 1:
            return self.g
        # the camera centres are gathered once per view assignment, not per step — but a rank whose view changed (the normal
        # training loop round-robins the views) must not form dL/dshN with last step's directions: ranks vote (4 bytes) on
        # "my centre differs from the cached one" and everybody re-gathers if anyone's did
        mine = campos_local.detach().to("cpu", torch.float32).reshape(3)
        changed = not getattr(self, "_cams_set", False) or not torch.equal(mine, self._cam_local)
        if getattr(self, "_cams_set", False):
            flag = torch.tensor([int(changed)], dtype=torch.int32, device=self.campos_all.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            changed = bool(int(flag.item()))
        if changed:
            self.set_cameras(campos_local)
        dist.all_gather_into_tensor(self.dsh0_all.view(-1), self.g.sh0.reshape(-1), group=self.group)
        dist.all_reduce(self.range_a, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(self.range_b, op=dist.ReduceOp.SUM, group=self.group)
        if self.g.shN.numel():
            self._accumulate(means, self.campos_all, self.dsh0_all, deg, self.g.shN)
        return self.g

    @staticmethod
    def wire_bytes_per_gaussian(world: int, sh_rest: int = 15) -> float:
        return (world - 1) * 12 + 2 * (world - 1) / world * 56

    @staticmethod
    def plain_wire_bytes_per_gaussian(world: int, sh_rest: int = 15) -> float:
        return 2 * (world - 1) / world * (56 + 12 * sh_rest)


class FusedGradientExchange:
    """The whole exchange as ONE kernel of ours over NVSwitch (csrc/collective.cu, dvs_coll_exchange_fused): a device-side
    cross-rank barrier (multimem.red on a symmetric counter), then a third of the CTAs sum the 44 B per Gaussian of
    quats | means, scales | opacities in the switch (multimem.ld_reduce / multimem.st) while all CTAs read every view's 12 B per
    Gaussian of dL/dsh0 straight from that rank's arena (NVLink peer loads) and form the summed dL/dshN and dL/dsh0 locally,
    and a second device-side barrier.  No NCCL call, no host synchronisation, no staging buffer, one launch.  Needs the gradient
    arena in symmetric memory with a multicast mapping (GradientReducer with backend nvls/auto provides it: pass its `flat` and
    handle) and one view per rank and step.

    With `DVS_FLAG_SKIP_SHN_GRAD` in the backward the per-view dL/dshN (180 B per Gaussian) is never written to HBM at all: the
    kernel writes the sum.
    """

    def __init__(self, grads, reducer: "GradientReducer", group=None, ctas: int = 0, reduce_ctas: int = 0):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _cabi
        if reducer._hdl is None or reducer.flat.data_ptr() != grads.flat.data_ptr():
            raise RuntimeError("FusedGradientExchange: the gradient arena must be the reducer's symmetric (NVLS) buffer")
        if grads.shN.numel() == 0:
            raise ValueError("FusedGradientExchange: no higher SH bands (degree 0), nothing to factor out")
        self.g, self.group = grads, group
        self._C, self._cabi, self._lib = C, _cabi, _cabi.load()
        hdl = reducer._hdl
        self.world, self.rank = hdl.world_size, hdl.rank
        if self.world > 16:
            raise ValueError("FusedGradientExchange: at most 16 ranks")
        dev = grads.flat.device
        N = grads.opacities.shape[0]
        self.N = N
        group_name = (group or dist.group.WORLD).group_name
        self.signal = symm_mem.empty(64, dtype=torch.int32, device=dev)
        self.signal.zero_()
        self._sh = symm_mem.rendezvous(self.signal, group_name)
        if not getattr(self._sh, "multicast_ptr", 0):
            raise RuntimeError("FusedGradientExchange: no NVLS multicast mapping for the signal buffer")
        peers = list(hdl.buffer_ptrs)
        if len(peers) != self.world or not all(peers):
            raise RuntimeError("FusedGradientExchange: the symmetric arena has no peer mappings")
        self.local_words = torch.zeros(64, dtype=torch.int32, device=dev)  # [0] grid counter, [1] tile counter, [32] status
        self.sh0_tmp = torch.zeros((3 * N + 3) // 4 * 4, dtype=torch.float32, device=dev)
        flat = grads.flat
        off = lambda t: (t.data_ptr() - flat.data_ptr()) // 4  # noqa: E731
        # arena order: quats | shN | means3D | scales | sh0 | opacities (rasterizer.GradBuffers.allocate)
        if not (off(grads.quats) < off(grads.shN) < off(grads.means3D) < off(grads.scales) < off(grads.sh0) < off(grads.opacities)):
            raise ValueError("FusedGradientExchange: unexpected gradient arena layout")
        a = _cabi.DvsCollFused()
        a.arena_mc, a.arena_local = hdl.multicast_ptr, flat.data_ptr()
        for r, ptr in enumerate(peers):
            a.arena_peers[r] = ptr
        a.sh0_tmp = self.sh0_tmp.data_ptr()
        a.signal_mc, a.signal_local = self._sh.multicast_ptr, self.signal.data_ptr()
        a.grid_counter, a.status = self.local_words.data_ptr(), self.local_words.data_ptr() + 128
        a.N, a.off_sh0, a.off_shN = N, off(grads.sh0), off(grads.shN)
        a.ranges[0][0], a.ranges[0][1] = off(grads.quats), off(grads.quats) + 4 * N
        a.ranges[1][0], a.ranges[1][1] = off(grads.means3D), off(grads.sh0)  # means3D | scales (+ alignment pad, zero everywhere)
        end_c = min((off(grads.opacities) + N + 3) // 4 * 4, flat.numel())   # the pad floats behind the last tensor are zero
        a.ranges[2][0], a.ranges[2][1] = off(grads.opacities), end_c
        for r in range(3):
            assert a.ranges[r][0] % 4 == 0 and a.ranges[r][1] % 4 == 0, "GradBuffers keeps every tensor 16-byte aligned"
        a.rank, a.world, a.sh_rest_alloc = self.rank, self.world, grads.shN.shape[1]
        a.ctas, a.reduce_ctas = ctas, reduce_ctas
        self._args = a
        self.launches = 0
        self.grid = self._lib.dvs_coll_exchange_fused_grid(ctas, self.world)
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # every rank's signal word is zero before anybody's first kernel touches it

    def set_cameras(self, campos_local: torch.Tensor):
        """All ranks' camera centres -> the kernel's by-value argument (gather once per view assignment; 12 bytes per rank)."""
        dev = self.g.flat.device
        allc = torch.zeros(self.world, 3, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allc.view(-1), campos_local.to(dev, torch.float32).reshape(3).contiguous(), group=self.group)
        host = allc.cpu().reshape(-1).tolist()
        for k, v in enumerate(host):
            self._args.campos[k] = v
        self._cam_local = campos_local.detach().to("cpu", torch.float32).reshape(3).clone()

    def exchange(self, means: torch.Tensor, campos_local: torch.Tensor, deg: int):
        if getattr(self, "_cam_local", None) is None:
            self.set_cameras(campos_local)
        elif not torch.equal(campos_local.detach().to("cpu", torch.float32).reshape(3), self._cam_local):
            # a changed view must be re-announced by EVERY rank in the same step (collective): callers that change views call
            # set_cameras() on all ranks; a silent local change would form dL/dshN with stale directions
            raise RuntimeError("FusedGradientExchange: this rank's camera changed; call set_cameras() on every rank first")
        a = self._args
        a.means, a.sh_degree, a.launch_index = means.data_ptr(), deg, self.launches
        st = torch.cuda.current_stream(self.g.flat.device).cuda_stream
        rc = self._lib.dvs_coll_exchange_fused(self._C.byref(a), self._C.c_void_p(st))
        if rc != 0:
            raise RuntimeError(f"dvs_coll_exchange_fused failed ({rc})")
        self.launches += 1
        return self.g

    def status(self) -> int:
        """Non-zero if a device-side barrier timed out (synchronises)."""
        return int(self.local_words[32].item())

    def wire_bytes_per_gaussian(self) -> float:
        return (self.world - 1) * 12 + 44 * (1 + 1 / self.world)
