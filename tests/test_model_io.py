"""SURVEY.md §8 row F2 — the model writers/readers (divshot_b200/csrc/model_io.cpp behind include/dvs_model_io.h)
against the REAL reference: external/tinygsplat + external/spz compiled unmodified into
oracle/_ref/libtinygsplat_ref.so (oracle/Makefile `ref`, shim oracle/ref_tinygsplat_shim.cpp).

Bar: every writer produces the SAME BYTES as the reference writer on the same input (tiny_gsplat.cpp:168-395,
994-1117, 1243-1272); every reader returns the SAME VALUES as the reference reader on the same file
(tiny_gsplat.cpp:632-816, 1119-1241), bit for bit.  Where the reference library is absent (no /root/reference and
no prebuilt oracle/_ref) the same comparisons run against the committed digests in tests/golden/model_io.json,
which tests/golden/make_model_io_golden.py froze from the reference writers."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libtinygsplat_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "model_io.json")
FORMATS = {1: "model.ply", 2: "model.splat", 3: "model.compressed.ply", 4: "model.dvsplat", 5: "model.spz",
           6: "model.reduced.ply"}
ROW = 59


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@pytest.fixture(scope="module")
def ours():
    from divshot_b200 import build
    lib = C.CDLL(os.environ.get("DVS_MODEL_IO_LIB") or build.build_gstrain())  # tools/fuzz_model_io.sh points this at a sanitizer build
    lib.dvs_model_read.restype = C.c_int64
    lib.dvs_model_read.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    lib.dvs_model_write.argtypes = [C.c_char_p, C.c_int, C.c_int64] + [C.c_void_p] * 7 + [C.c_uint32]
    lib.dvs_model_io_last_error.restype = C.c_char_p
    return lib


@pytest.fixture(scope="module")
def ref():
    if os.path.isdir("/root/reference/external/tinygsplat"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
    if not os.path.exists(REF_SO):
        pytest.skip("reference model-io library not built (no /root/reference, no prebuilt oracle/_ref)")
    lib = C.CDLL(REF_SO)
    lib.ref_load.restype = C.c_longlong
    lib.ref_load.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_longlong, C.c_void_p]
    lib.ref_save.argtypes = [C.c_int, C.c_char_p, C.c_longlong] + [C.c_void_p] * 7 + [C.c_int]
    return lib


def make_cloud(N, seed, degrees=None, spread=3.0):
    """Trainer-layout tensors of a plausible trained model (raw parameters)."""
    rng = np.random.default_rng(seed)
    c = dict(
        pos=(spread * rng.normal(size=(N, 3))).astype(np.float32),
        sh0=rng.normal(0, 1.2, size=(N, 3)).astype(np.float32),
        shn=rng.normal(0, 0.25, size=(N, 15, 3)).astype(np.float32),
        opac=rng.normal(0, 2.5, size=N).astype(np.float32),
        scale=rng.normal(-4, 1.5, size=(N, 3)).astype(np.float32),
        rot=rng.normal(size=(N, 4)).astype(np.float32),
        deg=None,
    )
    if degrees == "mixed":
        c["deg"] = rng.integers(0, 4, size=N).astype(np.uint8)
    return c


def write_ours(lib, fmt, path, c, flags=0):
    rc = lib.dvs_model_write(path.encode(), fmt, c["pos"].shape[0], _p(c["pos"]), _p(c["sh0"]), _p(c["shn"]),
                             _p(c["opac"]), _p(c["scale"]), _p(c["rot"]), _p(c["deg"]), flags)
    assert rc == 0, lib.dvs_model_io_last_error()


def write_ref(lib, fmt, path, c, aa=0):
    rc = lib.ref_save(fmt, path.encode(), c["pos"].shape[0], _p(c["pos"]), _p(c["sh0"]), _p(c["shn"]), _p(c["opac"]),
                      _p(c["scale"]), _p(c["rot"]), _p(c["deg"]), aa)
    assert rc == 0


def read_ours(lib, fmt, path):
    flags = C.c_uint32(0)
    n = lib.dvs_model_read(path.encode(), fmt, None, 0, C.byref(flags))
    assert n > 0, lib.dvs_model_io_last_error()
    rows = np.full((n, ROW), np.nan, np.float32)
    assert lib.dvs_model_read(path.encode(), fmt, _p(rows), n, C.byref(flags)) == n
    return rows, flags.value


def read_ref(lib, fmt, path, n):
    rows = np.full((n, ROW), np.nan, np.float32)
    aa = C.c_int(0)
    assert lib.ref_load(fmt, path.encode(), _p(rows), n, C.byref(aa)) == n
    return rows, aa.value


CASES = [  # (N, seed, degrees, antialiased): ragged last chunk, exactly one chunk, chunk multiple, tiny
    (1000, 1, None, 0), (256, 2, None, 1), (4096, 3, None, 0), (3, 4, None, 0), (20000, 5, None, 1),
]


@pytest.mark.parametrize("fmt", sorted(FORMATS))
@pytest.mark.parametrize("N,seed,degrees,aa", CASES)
def test_writer_bytes_equal_the_reference_writer(ours, ref, tmp_path, fmt, N, seed, degrees, aa):
    c = make_cloud(N, seed, degrees)
    a, b = str(tmp_path / ("ours_" + FORMATS[fmt])), str(tmp_path / ("ref_" + FORMATS[fmt]))
    write_ours(ours, fmt, a, c, flags=aa)
    write_ref(ref, fmt, b, c, aa)
    assert open(a, "rb").read() == open(b, "rb").read()


def test_dvsplat_with_mixed_sh_degrees_equals_the_reference(ours, ref, tmp_path):
    c = make_cloud(5000, 11, "mixed")
    a, b = str(tmp_path / "a.dvsplat"), str(tmp_path / "b.dvsplat")
    write_ours(ours, 4, a, c)
    write_ref(ref, 4, b, c)
    assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("N,seed", [(5000, 12), (3, 13), (257, 14)])
def test_reduced_ply_with_mixed_sh_degrees_equals_the_reference(ours, ref, tmp_path, N, seed):
    """Four per-degree vertex blocks (some possibly empty), rows of different widths; writer bytes, reader rows, and the
    reader on the half-float-position variant only the reference writer can produce (tiny_gsplat.cpp:398-630, 817-992)."""
    c = make_cloud(N, seed, "mixed")
    a, b, h = (str(tmp_path / n) for n in ("a.reduced.ply", "b.reduced.ply", "h.reduced.ply"))
    write_ours(ours, 6, a, c)
    write_ref(ref, 6, b, c)
    assert open(a, "rb").read() == open(b, "rb").read()
    mine, _ = read_ours(ours, 6, b)
    theirs, _ = read_ref(ref, 6, b, N)
    assert np.array_equal(mine.view(np.uint32), theirs.view(np.uint32))
    write_ref(ref, 7, h, c)  # format 7 of the shim: the same writer with halfFloat = true
    mine, _ = read_ours(ours, 6, h)
    theirs, _ = read_ref(ref, 7, h, N)
    assert np.array_equal(mine.view(np.uint32), theirs.view(np.uint32))


def test_reduced_ply_layout_and_the_fixed_sh_variant(ours, tmp_path):
    """Without the reference: rows are grouped by degree in input order; the default writer reproduces the reference's
    overlapping SH windows (quirk Q8: coefficient j = floats j..j+2 of the 45-float row), DVS_IO_REDUCED_SH_FIXED writes the
    coefficients themselves and then the round trip is lossless for every coefficient a degree stores."""
    c = make_cloud(900, 15, "mixed")
    order = np.concatenate([np.flatnonzero(c["deg"] == d) for d in range(4)])
    flat = c["shn"].reshape(-1, 45)
    for flags, window in ((0, lambda i, j: flat[i, j:j + 3]), (4, lambda i, j: flat[i, 3 * j:3 * j + 3])):
        p = str(tmp_path / f"m{flags}.reduced.ply")
        write_ours(ours, 6, p, c, flags=flags)
        r, _ = read_ours(ours, 0, p)  # DVS_FMT_AUTO picks the reduced reader from the name
        assert r.shape[0] == 900
        assert np.array_equal(r[:, :3], c["pos"][order]) and np.array_equal(r[:, 3:6], c["sh0"][order])
        assert np.array_equal(r[:, 51], c["opac"][order]) and np.array_equal(r[:, 52:55], c["scale"][order])
        assert np.array_equal(r[:, 55:], c["rot"][order])
        for row, i in enumerate(order[::37]):
            k = (int(c["deg"][i]) + 1) ** 2 - 1
            got = r[row * 37, 6:51].reshape(15, 3)
            assert all(np.array_equal(got[j], window(i, j)) for j in range(k)) and not got[k:].any()
    head = open(p, "rb").read().split(b"end_header\n")[0]
    assert head.count(b"element vertex") == 4 and b"comment generated by diverseshot" in head


def test_degenerate_extents_and_extreme_values_equal_the_reference(ours, ref, tmp_path):
    """A planar cloud (zero extent in z: NaN Morton cell in the reference), saturating colours / opacities / scales,
    quaternions with negative and tied largest components, duplicate positions (equal Morton codes)."""
    c = make_cloud(3000, 21)
    c["pos"][:, 2] = 1.5
    c["pos"][100:400] = c["pos"][100]
    c["sh0"][:50] = 9.0
    c["sh0"][50:100] = -9.0
    c["opac"][:30] = 40.0
    c["opac"][30:60] = -40.0
    c["scale"][:20] = 8.0
    c["scale"][20:40] = -12.0
    c["rot"][:10] = np.array([-1, 0, 0, 0], np.float32)
    c["rot"][10:20] = np.array([0.5, -0.5, 0.5, -0.5], np.float32)
    c["rot"][20:30] = np.array([0, 0, -3, 3], np.float32)
    c["shn"][:40] = 2.0
    c["shn"][40:80] = -2.0
    for fmt, name in FORMATS.items():
        a, b = str(tmp_path / ("a_" + name)), str(tmp_path / ("b_" + name))
        write_ours(ours, fmt, a, c)
        write_ref(ref, fmt, b, c)
        assert open(a, "rb").read() == open(b, "rb").read(), name


@pytest.mark.parametrize("fmt", sorted(FORMATS))
def test_reader_values_equal_the_reference_reader(ours, ref, tmp_path, fmt):
    c = make_cloud(2500, 31 + fmt)
    path = str(tmp_path / FORMATS[fmt])
    write_ref(ref, fmt, path, c, 1)
    mine, flags = read_ours(ours, fmt, path)
    theirs, aa = read_ref(ref, fmt, path, 2500)
    assert mine.shape == theirs.shape
    assert np.array_equal(mine.view(np.uint32), theirs.view(np.uint32)), np.argwhere(mine != theirs)[:10]
    if fmt in (1, 3, 5):
        assert flags == 1 and aa == 1


def test_ply_round_trip_is_lossless_and_layout_is_channel_major(ours, tmp_path):
    c = make_cloud(777, 41)
    path = str(tmp_path / "m.ply")
    write_ours(ours, 1, path, c)
    rows, flags = read_ours(ours, 1, path)
    assert flags == 0
    assert np.array_equal(rows[:, :3], c["pos"]) and np.array_equal(rows[:, 3:6], c["sh0"])
    assert np.array_equal(rows[:, 6:51].reshape(-1, 3, 15), c["shn"].transpose(0, 2, 1))  # tiny_gsplat.cpp:231-236
    assert np.array_equal(rows[:, 51], c["opac"]) and np.array_equal(rows[:, 52:55], c["scale"])
    assert np.array_equal(rows[:, 55:], c["rot"])


def test_lossy_formats_decode_close_to_the_input(ours, tmp_path):
    """Self-consistency without the reference: each quantised format decodes to within its quantisation step."""
    c = make_cloud(4000, 51, spread=1.0)
    unit = c["rot"] / np.linalg.norm(c["rot"], axis=1, keepdims=True)
    act = 1 / (1 + np.exp(-c["opac"].astype(np.float64)))
    # .splat (same order)
    p = str(tmp_path / "m.splat")
    write_ours(ours, 2, p, c)
    r, _ = read_ours(ours, 2, p)
    assert np.array_equal(r[:, :3], c["pos"]) and np.allclose(r[:, 52:55], c["scale"], atol=1e-5)
    assert np.abs(r[:, 55:] - unit).max() <= 1 / 128 + 1e-6
    # .spz (same order): 12 fractional bits, 1/16 log-scale steps, 9-bit quaternion components
    p = str(tmp_path / "m.spz")
    write_ours(ours, 5, p, c, flags=2)  # corrected SH indexing
    r, _ = read_ours(ours, 5, p)
    assert np.abs(r[:, :3] - c["pos"]).max() <= 0.5 / 4096 + 1e-7
    assert np.abs(r[:, 52:55] - np.clip(c["scale"], -10, 5.9375)).max() <= 1 / 32 + 1e-6
    sgn = np.sign(np.sum(r[:, 55:] * unit, axis=1, keepdims=True))
    assert np.abs(r[:, 55:] * sgn - unit).max() < 4e-3
    assert np.abs(1 / (1 + np.exp(-r[:, 51].astype(np.float64))) - act)[np.abs(c["opac"]) < 5].max() <= 0.5 / 255 + 1e-6
    assert np.abs(r[:, 6:51] - np.clip(c["shn"].reshape(-1, 45), -1, 0.99)).max() <= 8.5 / 128 + 1e-6  # half a 16/128 bucket + the 1/256 pre-rounding
    # compressed PLY / .dvsplat are Morton-ordered: compare as sets via the decoded positions
    for fmt, name, tol in ((3, "m.compressed.ply", 8.0 / 1023), (4, "m.dvsplat", 8.0 / 1023)):
        p = str(tmp_path / name)
        write_ours(ours, fmt, p, c)
        r, _ = read_ours(ours, fmt, p)
        assert r.shape[0] == 4000
        lo, hi = c["pos"].min(0), c["pos"].max(0)
        assert (r[:, :3] >= lo - 1e-4).all() and (r[:, :3] <= hi + 1e-4).all()
        # nearest input point of every decoded point is within the chunk quantisation step
        from scipy.spatial import cKDTree
        d, _ = cKDTree(c["pos"]).query(r[:, :3])
        assert d.max() < tol * np.linalg.norm(hi - lo)


def test_auto_format_dispatch_and_errors(ours, tmp_path):
    f = ours.dvs_model_format_from_path
    assert f(b"/x/a.ply") == 1 and f(b"a.compressed.ply") == 3 and f(b"a.splat") == 2
    assert f(b"a.dvsplat") == 4 and f(b"a.spz") == 5 and f(b"a.obj") == 0 and f(None) == 0
    assert f(b"a.reduced.ply") == 6 and f(b"a.reduced.compressed.ply") == 3  # ".compressed" is looked for first
    c = make_cloud(10, 1)
    rc = ours.dvs_model_write(str(tmp_path / "a.obj").encode(), 0, 10, _p(c["pos"]), _p(c["sh0"]), _p(c["shn"]),
                              _p(c["opac"]), _p(c["scale"]), _p(c["rot"]), None, 0)
    assert rc < 0 and b"unknown model format" in ours.dvs_model_io_last_error()
    rc = ours.dvs_model_write(b"/nonexistent_dir/a.ply", 0, 10, _p(c["pos"]), _p(c["sh0"]), _p(c["shn"]),
                              _p(c["opac"]), _p(c["scale"]), _p(c["rot"]), None, 0)
    assert rc < 0 and b"cannot write" in ours.dvs_model_io_last_error()
    assert ours.dvs_model_read(b"/nonexistent_dir/a.ply", 0, None, 0, None) < 0
    bad = tmp_path / "bad.spz"
    bad.write_bytes(b"not gzip")
    assert ours.dvs_model_read(str(bad).encode(), 0, None, 0, None) < 0
    trunc = tmp_path / "t.ply"
    write_ours(ours, 1, str(trunc), c)
    trunc.write_bytes(trunc.read_bytes()[:-100])
    assert ours.dvs_model_read(str(trunc).encode(), 0, None, 0, None) < 0
    assert b"truncated" in ours.dvs_model_io_last_error()


def test_plugin_writer_hook_uses_the_same_writers(ours, tmp_path):
    """gstrain_write_model (what save_splat_model calls) dispatches by extension to the same code."""
    c = make_cloud(300, 61)
    for fmt, name in FORMATS.items():
        a, b = str(tmp_path / ("hook_" + name)), str(tmp_path / ("api_" + name))
        ours.gstrain_write_model.argtypes = [C.c_char_p, C.c_longlong] + [C.c_void_p] * 6
        assert ours.gstrain_write_model(a.encode(), 300, _p(c["pos"]), _p(c["sh0"]), _p(c["shn"]), _p(c["opac"]),
                                        _p(c["scale"]), _p(c["rot"])) == 0
        write_ours(ours, fmt, b, c)
        assert open(a, "rb").read() == open(b, "rb").read()


def test_writers_match_the_committed_reference_digests(ours, tmp_path):
    """Travels to boxes without /root/reference: sha256 of the REFERENCE writers' files, frozen by
    tests/golden/make_model_io_golden.py, must equal the sha256 of our files on the same seeded clouds."""
    gold = json.load(open(GOLDEN))
    for case in gold["cases"]:
        c = make_cloud(case["N"], case["seed"], case["degrees"])
        for fmt_s, digest in case["sha256"].items():
            path = str(tmp_path / FORMATS[int(fmt_s)])
            write_ours(ours, int(fmt_s), path, c, flags=case["aa"])
            assert hashlib.sha256(open(path, "rb").read()).hexdigest() == digest, (case, fmt_s)


def test_readers_reject_forged_counts_and_damaged_files(ours, tmp_path):
    """Readers face files from anywhere (the viewer opens what the user drops on it): a header whose vertex count would wrap
    the size computation, truncations and random damage must end in an error or a bounded read, never outside the file
    (this test is also run under AddressSanitizer when the readers change: tools/fuzz_model_io.sh)."""
    c = make_cloud(300, 71, "mixed")
    rng = np.random.default_rng(5)
    for fmt, name in FORMATS.items():
        p = str(tmp_path / name)
        write_ours(ours, fmt, p, c)
        good = open(p, "rb").read()
        if fmt in (1, 3, 6):  # text headers: forge the count
            for forged in (b"18446744073709551615", b"4611686018427387904", b"99999999999"):
                bad = good.replace(b"element vertex 300", b"element vertex " + forged, 1) if fmt != 6 else \
                    good.replace(b"element vertex ", b"element vertex " + forged[:-1], 1)
                q = str(tmp_path / ("forged_" + name))
                open(q, "wb").write(bad)
                assert ours.dvs_model_read(q.encode(), fmt, None, 0, None) < 0, (name, forged)
                # the header ends the file WITHOUT a trailing newline: there is no payload position at all (the size checks
                # used to wrap and the row loop read past the buffer)
                cut = bad[:bad.index(b"end_header") + len(b"end_header")]
                open(q, "wb").write(cut)
                assert ours.dvs_model_read(q.encode(), fmt, None, 0, None) < 0, (name, "no newline after end_header")
                err = ours.dvs_model_io_last_error()
                assert b"truncated" in err or b"end_header" in err, err
                rows = np.zeros((4, ROW), np.float32)
                assert ours.dvs_model_read(q.encode(), fmt, _p(rows), 4, None) < 0
            cut = good[:good.index(b"end_header") + len(b"end_header")]
            open(q, "wb").write(cut)
            assert ours.dvs_model_read(q.encode(), fmt, None, 0, None) < 0, (name, "header only, no newline")
        for trial in range(40):
            b = bytearray(good)
            if trial % 2:
                b = b[:int(rng.integers(0, len(b)))]
            else:
                for _ in range(int(rng.integers(1, 12))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            q = str(tmp_path / ("damaged_" + name))
            open(q, "wb").write(bytes(b))
            n = ours.dvs_model_read(q.encode(), fmt, None, 0, None)
            if 0 < n <= 100000:
                rows = np.zeros((n, ROW), np.float32)
                assert ours.dvs_model_read(q.encode(), fmt, _p(rows), n, None) in (n, -1)
