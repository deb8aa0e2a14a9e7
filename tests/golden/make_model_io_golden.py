"""Freezes sha256 digests of the REFERENCE writers' output (oracle/_ref/libtinygsplat_ref.so = the unmodified
external/tinygsplat + external/spz of /root/reference) for seeded clouds -> tests/golden/model_io.json.
Run here (the reference tree is needed):  python tests/golden/make_model_io_golden.py"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

from test_model_io import FORMATS, REF_SO, make_cloud, write_ref  # noqa: E402

CASES = [dict(N=1000, seed=101, degrees=None, aa=0), dict(N=5000, seed=102, degrees="mixed", aa=1),
         dict(N=70000, seed=103, degrees=None, aa=0)]

if __name__ == "__main__":
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    ref = C.CDLL(REF_SO)
    ref.ref_save.argtypes = [C.c_int, C.c_char_p, C.c_longlong] + [C.c_void_p] * 7 + [C.c_int]
    out = {"source": "reference writers: external/tinygsplat/tiny_gsplat.cpp + external/spz/src/load-spz.cc", "cases": []}
    with tempfile.TemporaryDirectory() as d:
        for case in CASES:
            c = make_cloud(case["N"], case["seed"], case["degrees"])
            digests = {}
            for fmt, name in FORMATS.items():
                p = os.path.join(d, name)
                write_ref(ref, fmt, p, c, case["aa"])
                digests[str(fmt)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
            out["cases"].append({**case, "sha256": digests})
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "model_io.json"), "w"), indent=1)
    print("wrote tests/golden/model_io.json")
