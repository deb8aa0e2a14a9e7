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
    "aux_outputs.cu": [],
    "collective.cu": [],
    "sh_exchange.cu": [],
    "surfel.cu": [],
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


def _variant_flags():
    """A/B experiments only: DVS_NVCC_DEFINES="DVS_TIGHT_TILES DVS_RB_ROUND=128" adds -D macros to every kernel file (use
    with force=True into a scratch copy of the tree; the default build defines none)."""
    return [f"-D{d}" for d in os.environ.get("DVS_NVCC_DEFINES", "").split() if d]


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
            cmd = [nvcc(), "-ccbin", host_cxx(), *ARCH, *COMMON, *extra, *_variant_flags(), "-c", s, "-o", o]
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


def build_variant(name: str, defines=(), rev: str | None = None, files=None) -> str:
    """A/B experiments only: a second libdvsrast built with extra -D macros (and/or from the sources of an earlier git
    revision `rev`) -> divshot_b200/lib/variants/libdvsrast_<name>.so; tools/ab_bench.py loads it through DVS_RAST_LIB.
    The default build never depends on it."""
    vdir = os.path.join(ROOT, "build", "variants", name)
    vobj = os.path.join(vdir, "obj")
    os.makedirs(vobj, exist_ok=True)
    csrc, inc = CSRC, os.path.join(ROOT, "include")
    if rev:
        csrc, inc = os.path.join(vdir, "src", "csrc"), os.path.join(vdir, "src", "include")
        for d in (csrc, inc):
            shutil.rmtree(d, ignore_errors=True)
            os.makedirs(d)
        for sub, dst in (("divshot_b200/csrc", csrc), ("include", inc)):
            names = subprocess.check_output(["git", "ls-tree", "--name-only", f"{rev}:{sub}"], cwd=ROOT, text=True).split()
            for n in names:
                with open(os.path.join(dst, n), "wb") as f:
                    f.write(subprocess.check_output(["git", "show", f"{rev}:{sub}/{n}"], cwd=ROOT))
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", inc, "-I", csrc,
              "--expt-relaxed-constexpr"]
    procs, objs = [], []
    for src, extra in CU_SOURCES.items():
        s = os.path.join(csrc, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(vobj, src.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc(), "-ccbin", host_cxx(), *ARCH, *common, *extra, *[f"-D{d}" for d in defines], "-c", s, "-o", o]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(f"--- {name}: {src} ---\n{out}\n")
            raise RuntimeError(f"variant {name}: nvcc failed on {src}")
    vout = os.path.join(OUT, "variants")
    os.makedirs(vout, exist_ok=True)
    so = os.path.join(vout, f"libdvsrast_{name}.so")
    subprocess.check_call([nvcc(), "-ccbin", host_cxx(), *ARCH, "-shared", "-o", so, *objs, "-cudart", "static"])
    return so


def build_model_io(force: bool = False) -> str:
    """model_io.o: the host-only model writers/readers (include/dvs_model_io.h).  g++, -ffp-contract=off and the
    x86-64 baseline ISA (no FMA): its quantisers are literal operation sequences, byte-exact vs the reference."""
    os.makedirs(OBJ, exist_ok=True)
    src = os.path.join(CSRC, "model_io.cpp")
    obj = os.path.join(OBJ, "model_io.o")
    if force or _stale(obj, [src, os.path.join(ROOT, "include", "dvs_model_io.h")]):
        subprocess.check_call([host_cxx(), "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-fvisibility=hidden",
                               "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj])
    return obj


def build_gstrain(force: bool = False) -> str:
    """libgstrain.so: the trainer plugin (nine C symbols) + the rasterizer objects + the model writers/readers,
    one self-contained library."""
    build_rast(force)
    io_obj = build_model_io(force)
    src = os.path.join(CSRC, "gstrain.cu")
    hdr = os.path.join(ROOT, "include", "gaussian_trainer_scene.hpp")
    obj = os.path.join(OBJ, "gstrain.o")
    if force or _stale(obj, [src, hdr, os.path.join(ROOT, "include", "dvs_rast.h"),
                             os.path.join(ROOT, "include", "dvs_model_io.h"), os.path.join(ROOT, "include", "dvs_viewer_pack.h"),
                             os.path.join(CSRC, "densify.h")]):
        subprocess.check_call([nvcc(), "-ccbin", host_cxx(), *ARCH, *COMMON, "-c", src, "-o", obj])
    dsrc, dobj = os.path.join(CSRC, "densify.cu"), os.path.join(OBJ, "densify.o")  # trainer refinement step (F1)
    if force or _stale(dobj, [dsrc, os.path.join(CSRC, "densify.h"), os.path.join(CSRC, "densify_ops.h")]):
        subprocess.check_call([nvcc(), "-ccbin", host_cxx(), *ARCH, *COMMON, "-c", dsrc, "-o", dobj])
    # trainer -> viewer hand-off (F3): literal operation sequence, contraction off like preprocess_fwd.cu
    vsrc, vobj = os.path.join(CSRC, "viewer_pack.cu"), os.path.join(OBJ, "viewer_pack.o")
    if force or _stale(vobj, [vsrc, os.path.join(CSRC, "viewer_pack_ops.h"), os.path.join(ROOT, "include", "dvs_viewer_pack.h")]):
        subprocess.check_call([nvcc(), "-ccbin", host_cxx(), *ARCH, *COMMON, "-fmad=false", "-c", vsrc, "-o", vobj])
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in CU_SOURCES] + [obj, io_obj, dobj, vobj]
    so = os.path.join(OUT, "libgstrain.so")
    if force or _stale(so, objs):
        subprocess.check_call([nvcc(), "-ccbin", host_cxx(), *ARCH, "-shared", "-o", so, *objs, "-cudart", "static", "-lz"])
    return so


def build_torch_binding(force: bool = False) -> str:
    """libdvs_torch.so: torch::CustomClassHolder + autograd op over the C-ABI (g++ only, links libdvsrast.so)."""
    build_rast(force)
    src = os.path.join(CSRC, "torch_binding.cpp")
    so = os.path.join(OUT, "libdvs_torch.so")
    if not (force or _stale(so, [src, os.path.join(ROOT, "include", "dvs_rast.h")])):
        return so
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths()] + ["-I/usr/local/cuda/include", f"-I{os.path.join(ROOT, 'include')}"]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch.compiled_with_cxx11_abi())
    cmd = [host_cxx(), "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc, src, "-o", so,
           f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda", f"-L{OUT}", "-ldvsrast",
           f"-Wl,-rpath,{tlib}", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return so


def build_reference_cli(force: bool = False, app: str = "diverseshot-cli"):
    """Compile the reference's UNMODIFIED CLI sources (application/diverseshot-cli, or its older sibling
    application/splatx-cli) against include/gaussian_trainer_scene.hpp
    (recipe verified in SURVEY.md section 8-b).  Only possible where /root/reference exists; the binary goes to
    build/refcli/ (git-ignored, travels to the GPU box).  plugin.cpp has a g++-13 compile error on Linux
    (plugin.cpp:93,125: std::format with a runtime string) so the loader is provided by tools/plugin_shim.cpp
    implementing the same core/plugin.h interface."""
    ref = "/root/reference"
    out_dir = os.path.join(ROOT, "build", "refcli")
    exe = os.path.join(out_dir, app)
    if not os.path.isdir(ref):
        return exe if os.path.exists(exe) else None
    shim = os.path.join(ROOT, "tools", "plugin_shim.cpp")
    if not force and not _stale(exe, [shim, os.path.join(ROOT, "include", "gaussian_trainer_scene.hpp")]):
        return exe
    os.makedirs(out_dir, exist_ok=True)
    import glob
    inc = [os.path.join(ROOT, "include"), f"{ref}/application/{app}/source", f"{ref}/diverse/diverse_base/source",
           f"{ref}/external/CLI11/include", f"{ref}/external", f"{ref}/external/spdlog/include", f"{ref}/external/glm"]
    srcs = [f"{ref}/application/{app}/source/main.cpp", f"{ref}/application/{app}/source/gs_train.cpp",
            f"{ref}/diverse/diverse_base/source/utility/file_utils.cpp", f"{ref}/diverse/diverse_base/source/core/ds_log.cpp",
            shim] + sorted(glob.glob(f"{ref}/external/spdlog/src/*.cpp"))
    cmd = [host_cxx(), "-std=c++20", "-O1", "-w", "-DDS_PLATFORM_LINUX", "-DDS_PLATFORM_UNIX", "-DSPDLOG_COMPILED_LIB",
           *[f"-I{i}" for i in inc], *srcs, "-o", exe, "-ldl", "-lpthread"]
    subprocess.check_call(cmd)
    return exe


def build_driver(force: bool = False) -> str:
    """tools/gstrain_driver.cpp -> build/gstrain_driver (dlopens libgstrain.so like the reference CLI does)."""
    src = os.path.join(ROOT, "tools", "gstrain_driver.cpp")
    exe = os.path.join(ROOT, "build", "gstrain_driver")
    if force or _stale(exe, [src, os.path.join(ROOT, "include", "gaussian_trainer_scene.hpp")]):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call([host_cxx(), "-std=c++17", "-O1", f"-I{os.path.join(ROOT, 'include')}", src, "-o", exe, "-ldl"])
    return exe


def build_editor_probe(force: bool = False) -> str:
    """tools/editor_link_probe.cpp -> build/editor_link_probe, LINKED against libgstrain.so the way the reference editor
    links its trainer (class methods, not dlsym)."""
    src = os.path.join(ROOT, "tools", "editor_link_probe.cpp")
    exe = os.path.join(ROOT, "build", "editor_link_probe")
    so = build_gstrain(force)
    if force or _stale(exe, [src, so, os.path.join(ROOT, "include", "gaussian_trainer_scene.hpp")]):
        subprocess.check_call([host_cxx(), "-std=c++17", "-O1", f"-I{os.path.join(ROOT, 'include')}", src, "-o", exe,
                               f"-L{OUT}", "-lgstrain", f"-Wl,-rpath,{OUT}", "-Wl,-rpath,$ORIGIN/../divshot_b200/lib"])
    return exe


def build_editor_api_probe(force: bool = False):
    """tools/editor_api_probe.cpp -> build/editor_api_probe: the editor's call patterns on the trainer class, compiled
    with the REFERENCE's glm on the include path (so only where /root/reference exists; the GPU box uses the prebuilt
    binary that travels in build/)."""
    src = os.path.join(ROOT, "tools", "editor_api_probe.cpp")
    exe = os.path.join(ROOT, "build", "editor_api_probe")
    glm = "/root/reference/external/glm"
    if not os.path.isdir(glm):
        return exe if os.path.exists(exe) else None
    so = build_gstrain(force)
    if force or _stale(exe, [src, so, os.path.join(ROOT, "include", "gaussian_trainer_scene.hpp")]):
        subprocess.check_call([host_cxx(), "-std=c++20", "-O1", "-Wall", f"-I{os.path.join(ROOT, 'include')}", f"-I{glm}",
                               "-DGLM_FORCE_INTRINSICS", "-DGLM_FORCE_DEPTH_ZERO_TO_ONE", "-DGLM_FORCE_SWIZZLE",
                               src, "-o", exe, f"-L{OUT}", "-lgstrain", f"-Wl,-rpath,{OUT}", "-Wl,-rpath,$ORIGIN/../divshot_b200/lib"])
    return exe


def build_all(force: bool = False, verbose: bool = False, torch_binding: bool = True):
    libs = {"libdvsrast": build_rast(force, verbose), "libgstrain": build_gstrain(force),
            "gstrain_driver": build_driver(force), "editor_link_probe": build_editor_probe(force)}
    cli = build_reference_cli(force)
    if cli:
        libs["reference_cli"] = cli
    cli2 = build_reference_cli(force, app="splatx-cli")
    if cli2:
        libs["reference_splatx_cli"] = cli2
    probe = build_editor_api_probe(force)
    if probe:
        libs["editor_api_probe"] = probe
    if torch_binding:
        libs["libdvs_torch"] = build_torch_binding(force)
    return libs


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
