#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_nvls_allreduce.py — correctness of the NVLS multimem all-reduce
(csrc/collective.cu) against ncclAllReduce on the same data, plus timing of both."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from divshot_b200.dp import GradientReducer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
numel = 59_000_000
red = GradientReducer(numel, dev, backend="nvls")
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
x = torch.randn(red.flat.numel(), device=dev, generator=g)
ref = x.clone()
dist.all_reduce(ref)
red.flat.copy_(x)
torch.cuda.synchronize(); dist.barrier()
red.all_reduce()
torch.cuda.synchronize()
err = float((red.flat - ref).abs().max() / ref.abs().max())
t_nvls = red._time(dev, reps=10)
sweep = {}
for ctas in (148, 296, 592, 1184):
    red.ctas = ctas
    sweep[ctas] = round(red._time(dev, reps=10), 3)
red.ctas = 0
if rank == 0:
    print("nvls ms by CTA count:", sweep)
red.backend = "nccl"
t_nccl = red._time(dev, reps=10)
if rank == 0:
    print(f"world {world}: max rel diff nvls vs nccl {err:.2e}; nvls {t_nvls:.3f} ms, nccl {t_nccl:.3f} ms "
          f"({red.flat.numel() * 4 / 1e6:.0f} MB)")
assert err < 1e-5, err
dist.destroy_process_group()
