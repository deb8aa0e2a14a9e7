"""N>1 host logic on CPU (gloo, world_size 2): view partition + single all-reduce of the gradient arena."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, N, KR, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from divshot_b200.dp import allreduce_gradients, views_for_rank
    from divshot_b200.rasterizer import GradBuffers
    g = GradBuffers.allocate(N, KR, torch.device("cpu"))
    # every view of the batch contributes a known pattern; this rank accumulates its own views
    mine = views_for_rank(8, rank, world)
    for v in mine:
        g.means3D += (v + 1); g.shN += 0.5 * (v + 1); g.opacities += 2.0 * (v + 1); g.quats -= (v + 1)
    allreduce_gradients(g.flat)
    tot = sum(range(1, 9))
    ok = (torch.allclose(g.means3D, torch.full_like(g.means3D, tot)) and torch.allclose(g.shN, torch.full_like(g.shN, 0.5 * tot))
          and torch.allclose(g.opacities, torch.full_like(g.opacities, 2.0 * tot)) and torch.allclose(g.quats, torch.full_like(g.quats, -tot))
          and float(g.scales.abs().sum()) == 0.0 and float(g.sh0.abs().sum()) == 0.0)
    q.put((rank, mine, bool(ok), g.flat.numel()))
    dist.destroy_process_group()


def test_view_partition_and_single_allreduce_world2():
    world, N, KR = 2, 1000, 15
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, KR, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]  # disjoint cover of the batch
    assert all(r[2] for r in res)
    assert res[0][3] >= N * (11 + 3 * (KR + 1))  # one arena holds all 11+3K floats per Gaussian


def test_gradbuffer_views_are_16_byte_aligned_and_disjoint():
    from divshot_b200.rasterizer import GradBuffers
    g = GradBuffers.allocate(1001, 15, torch.device("cpu"))
    base = g.flat.data_ptr()
    spans = []
    for name in ("means3D", "scales", "quats", "opacities", "sh0", "shN"):
        t = getattr(g, name)
        assert (t.data_ptr() - base) % 16 == 0, name
        spans.append((t.data_ptr() - base, t.data_ptr() - base + t.numel() * 4))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
