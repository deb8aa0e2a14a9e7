"""SURVEY.md §8(e) — the factored multi-GPU exchange: instead of all-reducing dL/dshN (180 of the 236 gradient bytes per
Gaussian), every rank all-gathers each view's dL/dsh0 (12 B per Gaussian) and forms sum_v B(dir_v) (x) dL/dsh0_v / SH_C0
itself (divshot_b200/csrc/sh_grad_ops.h, sh_exchange.cu; divshot_b200/dp.py FactoredGradientExchange).
CPU tier: the arithmetic (host build) against the oracle's per-view dL/dshN summed over views; the exchange logic with
gloo at world_size 2.  GPU tier: the kernel against the host build on a B200 (marker `gpu`)."""
import ctypes as C
import os
import socket
import subprocess

import numpy as np
import pytest
import torch

from divshot_b200.scenes import make_scene
from oracle import oracle as orc
from util import assert_close, orc_cam, scene_arrays

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ops():
    src = os.path.join(ROOT, "tests", "native", "sh_grad_host.cpp")
    hdr = os.path.join(ROOT, "divshot_b200", "csrc", "sh_grad_ops.h")
    out = os.path.join(ROOT, "build", "test_sh_grad_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-I", os.path.dirname(hdr), src, "-o", out])
    L = C.CDLL(out)
    L.t_sh_grad_from_dsh0.argtypes = [C.c_void_p] * 3 + [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.t_sh_exchange_kernel_emulation.argtypes = [C.c_void_p] * 3 + [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def host_sh_grad(means, campos, dsh0_all, deg, KR):
    V, N = dsh0_all.shape[:2]
    out = np.zeros((N, KR, 3), np.float32)
    _ops().t_sh_grad_from_dsh0(_p(np.ascontiguousarray(means, np.float32)), _p(np.ascontiguousarray(campos, np.float32)),
                               _p(np.ascontiguousarray(dsh0_all, np.float32)), N, V, deg, KR, _p(out))
    return out


def _per_view_grads(sc, deg):
    arrays = scene_arrays(sc)
    outs = []
    for v, cam in enumerate(sc.cameras):
        oc = orc_cam(cam, deg)
        f = orc.forward(oc, *arrays, threads=1)
        outs.append(orc.backward(oc, f, *arrays, sc.dL_dpix[v], threads=1))
    return outs


@pytest.mark.parametrize("deg,views", [(3, 4), (1, 3), (2, 2)])
def test_factored_sh_gradient_equals_the_sum_of_the_per_view_gradients(deg, views):
    sc = make_scene(N=3000, width=96, height=64, sh_degree=deg, views=views, seed=31 + deg)
    sc.log_scales += 0.8
    sc.sh0 -= 0.6  # push some colours below zero: the clamp mask must travel inside dL/dsh0
    b = _per_view_grads(sc, deg)
    KR = sc.shN.shape[1]
    campos = np.stack([c.campos for c in sc.cameras])
    dsh0_all = np.stack([x.dL_dsh0 for x in b])
    want = sum(x.dL_dshN.astype(np.float64) for x in b)
    got = host_sh_grad(sc.means3D, campos, dsh0_all, deg, KR)
    assert np.abs(want).max() > 0 and (dsh0_all == 0).all(axis=2).mean() > 0.05
    assert_close(got, want, 2e-5, "factored dL_dshN")
    K1 = (deg + 1) ** 2 - 1
    assert not got[:, K1:].any(), "bands above the active degree stay zero"


@pytest.mark.parametrize("N,V,deg,KR,grid,misalign", [(1000, 3, 3, 15, 4, 0), (128, 2, 2, 15, 1, 0), (129, 2, 1, 3, 7, 0), (1, 1, 3, 15, 2, 0),
                                                      (777, 8, 3, 15, 3, 1), (5000, 2, 0, 15, 5, 0), (300, 2, 1, 5, 2, 0)])
def test_kernel_indexing_thread_by_thread(N, V, deg, KR, grid, misalign):
    """The two phase functions the CUDA kernel calls between its barriers, run for every (CTA, thread) on the host with
    the kernel's own tile loop: partial last tile, vector / scalar store paths, rows narrower than 45 words, more CTAs than
    tiles, an output that is not 16-byte aligned.  Must equal the plain per-Gaussian evaluation and touch nothing else."""
    rng = np.random.default_rng(N * 7 + V)
    means = rng.normal(0, 3, (N, 3)).astype(np.float32)
    campos = rng.normal(0, 1, (V, 3)).astype(np.float32)
    dsh0 = rng.normal(0, 1, (V, N, 3)).astype(np.float32)
    dsh0[rng.random((V, N)) < 0.3] = 0
    buf = np.full(N * KR * 3 + 8, 777.0, np.float32)
    out = buf[4 + misalign:4 + misalign + N * KR * 3]
    _ops().t_sh_exchange_kernel_emulation(_p(means), _p(campos), _p(dsh0), N, V, deg, KR, C.c_void_p(out.ctypes.data), grid)
    want = host_sh_grad(means, campos, dsh0, deg, KR)
    assert np.array_equal(out.reshape(N, KR, 3), want)
    assert (buf[:4 + misalign] == 777.0).all() and (buf[4 + misalign + N * KR * 3:] == 777.0).all(), "wrote outside its rows"


# ------------------------------------------------------------------------------------------------ gloo, world_size 2
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from divshot_b200.dp import FactoredGradientExchange
    from divshot_b200.rasterizer import GradBuffers
    deg, N = 2, 1500
    sc = make_scene(N=N, width=64, height=48, sh_degree=deg, views=world, seed=77)
    sc.log_scales += 0.8
    KR = sc.shN.shape[1]
    grads = _per_view_grads(sc, deg)  # every rank can compute every view here; it only USES its own as its local gradient
    mine = grads[rank]
    g = GradBuffers.allocate(N, KR, torch.device("cpu"))
    g.means3D.copy_(torch.from_numpy(mine.dL_dmeans3D)); g.scales.copy_(torch.from_numpy(mine.dL_dscales))
    g.quats.copy_(torch.from_numpy(mine.dL_dquats)); g.opacities.copy_(torch.from_numpy(mine.dL_dopacities))
    g.sh0.copy_(torch.from_numpy(mine.dL_dsh0)); g.shN.copy_(torch.from_numpy(mine.dL_dshN))

    def accumulate(means, campos_all, dsh0_all, deg_, out_shN):  # stands in for the CUDA kernel of sh_exchange.cu
        out_shN.copy_(torch.from_numpy(host_sh_grad(means.numpy(), campos_all.numpy(), dsh0_all.numpy(), deg_, out_shN.shape[1])))

    ex = FactoredGradientExchange(g, accumulate=accumulate)
    ex.exchange(torch.from_numpy(sc.means3D), torch.from_numpy(np.asarray(sc.cameras[rank].campos, np.float32)), deg)
    tot = {k: sum(getattr(x, "dL_d" + k).astype(np.float64) for x in grads) for k in ("means3D", "scales", "quats", "opacities", "sh0", "shN")}
    ok = all(np.allclose(getattr(g, k).numpy().reshape(tot[k].shape), tot[k], rtol=2e-4, atol=2e-5 * np.abs(tot[k]).max()) for k in tot)
    q.put((rank, bool(ok), ex.wire_bytes_per_gaussian(world), ex.plain_wire_bytes_per_gaussian(world)))
    dist.destroy_process_group()


def test_factored_exchange_world2_gloo():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(r[1] for r in res), res
    assert res[0][2] < res[0][3]  # fewer bytes on the wire than the plain all-reduce, already at 2 ranks


@pytest.mark.gpu
@pytest.mark.parametrize("N,V,deg,KR", [(5000, 8, 3, 15), (12345, 2, 2, 15), (129, 3, 1, 3), (1000, 1, 0, 0), (70000, 4, 3, 15)])
def test_cuda_sh_accumulation_matches_the_host_build(N, V, deg, KR):
    from divshot_b200 import _cabi
    rng = np.random.default_rng(N + V)
    means = rng.normal(0, 3, (N, 3)).astype(np.float32)
    campos = rng.normal(0, 1, (V, 3)).astype(np.float32)
    dsh0 = rng.normal(0, 1, (V, N, 3)).astype(np.float32)
    dsh0[rng.random((V, N)) < 0.3] = 0  # views that do not see a Gaussian
    dev = torch.device("cuda", 0)
    t_means, t_d = torch.from_numpy(means).to(dev), torch.from_numpy(dsh0).to(dev)
    out = torch.full((N, max(KR, 1), 3), float("nan"), device=dev)
    lib = _cabi.load()
    rc = lib.dvs_coll_sh_grad_from_dsh0(t_means.data_ptr(), campos.ctypes.data, t_d.data_ptr(), N, V, deg, KR, out.data_ptr(),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    if KR == 0:
        return
    want = host_sh_grad(means, campos, dsh0, deg, KR)
    got = out.cpu().numpy()[:, :KR]
    assert np.isfinite(got).all()
    assert_close(got, want, 1e-5, "dL_dshN")


@pytest.mark.gpu
def test_factored_exchange_refuses_cpu_tensors():
    from divshot_b200.dp import FactoredGradientExchange
    from divshot_b200.rasterizer import GradBuffers
    g = GradBuffers.allocate(10, 15, torch.device("cpu"))
    ex = FactoredGradientExchange(g)
    with pytest.raises(RuntimeError):
        ex._accumulate_cuda(torch.zeros(10, 3), torch.zeros(1, 3), torch.zeros(1, 10, 3), 3, g.shN)


# ---- bench.py's collective decision (choose_exchange) on gloo: CUDA timing primitives are stubbed, the logic is real
class _FakeEvent:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def _choose_worker(rank, world, port, broken, q):
    import types

    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    import bench
    from divshot_b200.dp import FactoredGradientExchange
    from divshot_b200.rasterizer import GradBuffers
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = _FakeEvent
    deg, N = 2, 1200
    sc = make_scene(N=N, width=64, height=48, sh_degree=deg, views=world, seed=78)
    sc.log_scales += 0.8
    KR = sc.shN.shape[1]
    per_view = _per_view_grads(sc, deg)
    mine = per_view[rank]
    g = GradBuffers.allocate(N, KR, torch.device("cpu"))

    def load_local():
        g.means3D.copy_(torch.from_numpy(mine.dL_dmeans3D)); g.scales.copy_(torch.from_numpy(mine.dL_dscales))
        g.quats.copy_(torch.from_numpy(mine.dL_dquats)); g.opacities.copy_(torch.from_numpy(mine.dL_dopacities))
        g.sh0.copy_(torch.from_numpy(mine.dL_dsh0)); g.shN.copy_(torch.from_numpy(mine.dL_dshN))

    def accumulate(means, campos_all, dsh0_all, deg_, out_shN):
        out_shN.copy_(torch.from_numpy(host_sh_grad(means.numpy(), campos_all.numpy(), dsh0_all.numpy(), deg_, out_shN.shape[1])))
        if broken and rank == 1:  # only ONE rank computes garbage: the vote must still be unanimous
            out_shN.mul_(1.5)

    reducer = types.SimpleNamespace(flat=g.flat, backend="nccl", note="", all_reduce=lambda: dist.all_reduce(g.flat))
    fx = FactoredGradientExchange(g, accumulate=accumulate)
    campos = torch.from_numpy(np.asarray(sc.cameras[rank].campos, np.float32))
    fx.set_cameras(campos)
    load_local()
    args = types.SimpleNamespace(allreduce="auto")
    fn = bench.choose_exchange(args, reducer, fx, g, {"means3D": torch.from_numpy(sc.means3D)}, campos, deg, torch.device("cpu"), dist, torch)
    load_local()
    fn()
    tot = sum(x.dL_dshN.astype(np.float64) for x in per_view)
    tot_m = sum(x.dL_dmeans3D.astype(np.float64) for x in per_view)
    ok = (np.allclose(g.shN.numpy(), tot, rtol=2e-4, atol=2e-5 * np.abs(tot).max())
          and np.allclose(g.means3D.numpy(), tot_m, rtol=2e-4, atol=2e-5 * np.abs(tot_m).max()))
    q.put((rank, bool(ok), reducer.backend, reducer.note))
    dist.destroy_process_group()


@pytest.mark.parametrize("broken", [False, True])
def test_bench_adopts_the_factored_exchange_only_after_a_unanimous_self_check(broken):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_choose_worker, args=(r, world, port, broken, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(r[1] for r in res), res                       # whichever exchange was chosen sums correctly
    assert res[0][2] == res[1][2], "ranks must agree on the exchange"
    if broken:
        assert res[0][2] == "nccl" and all("rejected" in r[3] for r in res)
    else:
        assert all(("factored exchange" in r[3]) for r in res)  # timed against the plain one, either may win on gloo
