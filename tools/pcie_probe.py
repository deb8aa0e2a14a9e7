#!/usr/bin/env python
"""Host <-> device copy bandwidth per rank, alone and with all ranks copying at once (run under torch.distributed.run).
Explains the e2e numbers of bench.py at N > 1: every rank moves 19.2 MB in and 19.2 MB out per step through the host.
  python -m torch.distributed.run --nproc-per-node 8 ... tools/pcie_probe.py [--bind]
--bind: pin the process (and therefore its first-touch pinned allocations) to the CPUs of the GPU's NUMA node first."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--mb", type=int, default=19)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from divshot_b200.hostbind import bind_to_gpu_numa_node
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    note = bind_to_gpu_numa_node(local) if a.bind else "not bound"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.mb * 1000 * 1000 // 4
    h_in, h_out = torch.randn(n).pin_memory(), torch.empty(n).pin_memory()
    d_in, d_out = torch.empty(n, device=dev), torch.randn(n, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(mode, reps=20):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        return n * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9  # GB/s per direction

    res = {"rank": rank, "bind": note}
    for mode in ("h2d", "d2h", "both"):
        run(mode, 3)
        res[mode + "_all_ranks_GBps"] = round(run(mode), 1)
    if world > 1:  # one rank at a time
        for r in range(world):
            dist.barrier()
            if r == rank:
                res["both_alone_GBps"] = round(_alone(torch, n, h_in, h_out, d_in, d_out, s1, s2), 1)
            dist.barrier()
        rows = [None] * world
        dist.all_gather_object(rows, res)
    else:
        rows = [res]
    if rank == 0:
        print(json.dumps(rows))
    if world > 1:
        dist.destroy_process_group()


def _alone(torch, n, h_in, h_out, d_in, d_out, s1, s2, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return n * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


if __name__ == "__main__":
    main()
