"""Host-side placement for the end-to-end path: a rank's pinned host buffers should live on the NUMA node its GPU hangs off,
otherwise every H2D / D2H crosses the inter-socket link (8 ranks x 38 MB per step).  Linux first-touch policy: pin the
process to that node's CPUs BEFORE it allocates; nothing else is changed (no numactl / libnuma in the image)."""
from __future__ import annotations

import os
import subprocess


def _cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(gpu_index: int) -> int | None:
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bdf:
            return None
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:  # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_numa_node(gpu_index: int) -> str:
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with its current affinity).  Returns a note."""
    node = gpu_numa_node(gpu_index)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return "numa node of the GPU unknown: not bound"
    try:
        cpus = _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & os.sched_getaffinity(0)
        if not cpus:
            return f"numa node {node}: no allowed CPUs there, not bound"
        os.sched_setaffinity(0, cpus)
        return f"bound to numa node {node} ({len(cpus)} CPUs)"
    except Exception as e:  # noqa: BLE001
        return f"numa node {node}: binding failed ({type(e).__name__})"
