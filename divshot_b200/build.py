"""In-tree build of the native libraries (explicit nvcc / g++; no JIT cache, no pip install).

  divshot_b200/lib/libdvsrast.so   CUDA kernels + the C-ABI of include/dvs_rast.h (no torch dependency)

Everything is compiled for sm_100a only (`-gencode arch=compute_100a,code=sm_100a -lineinfo`).
preprocess_fwd.cu is compiled with -fmad=false: its arithmetic is a literal operation sequence
(bit-exact radii / tile rects / depth keys).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
OBJ = os.path.join(ROOT, "build", "obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", os.path.join(ROOT, "include"),
          "-I", CSRC, "--expt-relaxed-constexpr"]
CU_SOURCES = {
    "preprocess_fwd.cu": ["-fmad=false"],
    "preprocess_bwd.cu": [],
    "binning.cu": [],
    "render_fwd.cu": [],
    "render_bwd.cu": [],
    "api.cu": [],
}


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def host_cxx() -> str:
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_rast(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "dvs_rast.h"))
    objs = []
    procs = []
    for src, extra in CU_SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc(), "-ccbin", host_cxx(), *ARCH, *COMMON, *extra, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    so = os.path.join(OUT, "libdvsrast.so")
    if force or _stale(so, objs):
        cmd = [nvcc(), "-ccbin", host_cxx(), *ARCH, "-shared", "-o", so, *objs, "-cudart", "static"]
        subprocess.check_call(cmd)
    return so


def build_all(force: bool = False, verbose: bool = False):
    return {"libdvsrast": build_rast(force, verbose)}


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
