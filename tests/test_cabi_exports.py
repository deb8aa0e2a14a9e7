"""The C-ABI library loads without a GPU and exports every symbol include/dvs_rast.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dvs_rast.h")).read()
    return sorted(set(re.findall(r"DVS_API\s+[\w\s\*]+?\b(dvs_(?:rast|coll)_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for n in ("dvs_rast_create", "dvs_rast_destroy", "dvs_rast_forward", "dvs_rast_backward", "dvs_rast_step_host",
              "dvs_rast_last_error", "dvs_rast_get_stats", "dvs_rast_debug_read", "dvs_rast_reserve"):
        assert n in names


def test_library_exports_every_declared_symbol():
    from divshot_b200 import _cabi, build
    build.build_rast()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for n in _declared():
        assert hasattr(lib, n), f"libdvsrast.so does not export {n}"
    assert set(_cabi.EXPORTS + _cabi.COLL_EXPORTS) == set(_declared())
    lib.dvs_rast_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.dvs_rast_version()


def test_plugin_library_exports_every_model_io_symbol():
    """include/dvs_model_io.h is exported by libgstrain.so, which loads without a GPU."""
    from divshot_b200 import build
    src = open(os.path.join(ROOT, "include", "dvs_model_io.h")).read()
    names = sorted(set(re.findall(r"DVS_API\s+[\w\s\*]+?\b(dvs_model_\w+)\s*\(", src)))
    assert names == ["dvs_model_format_from_path", "dvs_model_io_last_error", "dvs_model_read", "dvs_model_write"]
    lib = ctypes.CDLL(build.build_gstrain())
    for n in names:
        assert hasattr(lib, n), f"libgstrain.so does not export {n}"


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from divshot_b200.rasterizer import Rasterizer, RasterizerError
    with pytest.raises(RasterizerError):
        Rasterizer(0)
    # and the raw C-ABI refuses too (no silent CPU path)
    from divshot_b200 import _cabi
    h = ctypes.c_void_p()
    assert _cabi.load().dvs_rast_create(0, ctypes.byref(h)) != 0


def test_struct_layouts_match_header():
    from divshot_b200 import _cabi
    assert ctypes.sizeof(_cabi.DvsCamera) == 4 * (16 + 16 + 3 + 2 + 2 + 3 + 1 + 3)
    assert ctypes.sizeof(_cabi.DvsParams) == 6 * 8 and ctypes.sizeof(_cabi.DvsGrads) == 8 * 8
    assert ctypes.sizeof(_cabi.DvsStats) == 5 * 8 + 4 * 4 + 8  # ... overflow, reserved_, num_list_entries
    assert _cabi.DvsStats.num_list_entries.offset == 56 and _cabi.DvsStats.overflow.offset == 48
    # dvs_coll_fused (static_assert'ed to the same numbers in csrc/collective.cu)
    assert ctypes.sizeof(_cabi.DvsCollFused) == 488 and _cabi.DvsCollFused.sh0_tmp.offset == 144
    assert _cabi.DvsCollFused.N.offset == 384 and _cabi.DvsCollFused.rank.offset == 464


def test_new_entry_points_validate_their_arguments_without_a_gpu():
    """Argument checks of the entry points added for rows F3 / F4 / §8(e) run before any CUDA call."""
    from divshot_b200 import _cabi, build
    L = _cabi.load()
    E_INVALID = 1
    lib_h = open(os.path.join(ROOT, "include", "dvs_rast.h")).read()
    assert re.search(r"#define\s+DVS_E_INVALID\s+1\b", lib_h) or "DVS_E_INVALID" in lib_h
    assert L.dvs_rast_forward_aux(None, None, None, None, None) != 0
    assert L.dvs_rast_backward_aux(None, None, None, None, None, None, 0, None) != 0
    # the SH exchange kernel: bad view counts / degrees / row widths are refused, an empty problem is a no-op
    f = L.dvs_coll_sh_grad_from_dsh0
    assert f(None, None, None, 10, 0, 3, 15, None, None) != 0       # no views
    assert f(None, None, None, 10, 65, 3, 15, None, None) != 0      # more than 64 views
    assert f(None, None, None, 10, 2, 4, 15, None, None) != 0       # degree > 3
    assert f(None, None, None, 10, 2, 3, 8, None, None) != 0        # rows too narrow for the degree
    assert f(None, None, None, 10, 2, 3, 15, None, None) != 0       # null pointers
    assert f(None, None, None, 0, 2, 3, 15, None, None) == 0        # N = 0
    g = ctypes.CDLL(build.build_gstrain())
    g.dvs_viewer_pack.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int64] + [ctypes.c_void_p] * 5
    assert g.dvs_viewer_pack(None, None, None, None, None, None, -1, None, None, None, None, None) != 0
    assert g.dvs_viewer_pack(None, None, None, None, None, None, 5, None, None, None, None, None) != 0  # no bbox buffer
    import torch
    if not torch.cuda.is_available():  # and without a device the launch itself fails loudly (no CPU path)
        bb = (ctypes.c_uint32 * 6)()
        assert g.dvs_viewer_pack(None, None, None, None, None, None, 0, None, None, None, bb, None) != 0
