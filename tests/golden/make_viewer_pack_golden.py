"""Freezes sha256 digests of the REFERENCE quantiser's output (oracle/_ref/libviewerpack_ref.so, i.e. the lines of
diverse/source/assets/gaussian_model.cpp compiled unmodified) for the models of tests/viewer_pack_util.make_model, so
boxes without /root/reference can still check the product's bytes.  Run in the build container:
    python tests/golden/make_viewer_pack_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import viewer_pack_util as u  # noqa: E402

CASES = [(20000, 1, 3), (5000, 2, 1), (333, 3, 0), (129, 5, 2), (128, 6, 3), (127, 7, 3), (1, 4, 3), (200003, 8, 3)]

if __name__ == "__main__":
    assert os.path.exists(u.REF_SO), "build oracle/_ref first: make -C oracle ref"
    out = {f"N{n}_seed{s}_deg{d}": u.digest(*u.pack_with_reference(u.make_model(n, s, d))) for n, s, d in CASES}
    json.dump({"source": "oracle/_ref/libviewerpack_ref.so (gaussian_model.cpp:14-22,126-128,130-211,292-298)", "sha256": out},
              open(os.path.join(HERE, "viewer_pack.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
