"""SURVEY.md §8 row F1 — divshot_b200/csrc/densify.cu executed on the CPU: the file is compiled for the host by
tests/native/densify_emul.cpp (kernel bodies as serial loops, the two CUB primitives as plain loops, the CUDA runtime
replaced by tests/native/cuda_host_shim.h), and driven through the same dvs_densify_test_* hooks and the same test
bodies (tests/densify_cases.py) as the GPU tests.  What this cannot see: launch geometry, the CUB calls, and
device libm rounding — those are what tests/test_zz_gpu_densify.py is for."""
import ctypes as C
import os
import subprocess

import pytest

import densify_cases as dc
from test_densify_ops import ops  # noqa: F401  (host build of densify_ops.h, used to predict noise / split samples)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def be():
    src = os.path.join(ROOT, "tests", "native", "densify_emul.cpp")
    csrc = os.path.join(ROOT, "divshot_b200", "csrc")
    deps = [src, os.path.join(ROOT, "tests", "native", "cuda_host_shim.h")] + [os.path.join(csrc, f) for f in ("densify.cu", "densify.h", "densify_ops.h")]
    out = os.path.join(ROOT, "build", "test_densify_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-I", os.path.dirname(src),
                               "-I", csrc, src, "-o", out])
    return dc.host_backend(C.CDLL(out))


@pytest.mark.parametrize("case", dc.CASES_PLAIN, ids=lambda c: c.__name__)
def test_refinement_step_on_the_host_build(be, case):
    case(be)


@pytest.mark.parametrize("case", dc.CASES_WITH_OPS, ids=lambda c: c.__name__)
def test_refinement_step_on_the_host_build_vs_per_element_ops(be, ops, case):  # noqa: F811
    case(be, ops)
