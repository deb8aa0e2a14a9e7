"""Shared test helpers: oracle camera from a scenes.Camera, tolerance metric."""
import numpy as np

from oracle import oracle as orc


def orc_cam(cam, sh_degree, sh_rest_alloc=None, flags=0):
    return orc.make_camera(cam.view, cam.proj, cam.campos, cam.tanfovx, cam.tanfovy, cam.width, cam.height,
                           cam.bg, cam.scale_modifier, sh_degree, sh_rest_alloc, flags)


def scene_arrays(sc):
    return (sc.means3D, sc.log_scales, sc.quats, sc.logit_opac, sc.sh0, sc.shN)


def rel_err(a, b):
    """Tolerance metric used for every floating-point parity claim in this repo:
    max |a-b| / (|b| + floor) with floor = the RMS of b (so that elements that are cancellation
    residue of much larger terms are judged on the tensor's scale, not on their own)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if b.size == 0:
        return 0.0
    floor = np.sqrt(np.mean(b * b)) + 1e-30
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"


def elem_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    floor = np.sqrt(np.mean(b * b)) + 1e-30
    return np.abs(a - b) / (np.abs(b) + floor)


def assert_close_robust(a, b, tol, what="", frac=0.999, loose=50.0):
    """Parity check that tolerates threshold ties.  The compositor takes hard decisions (alpha >= 1/255,
    T >= 1e-4) on values that differ in the last bits between two correct implementations (exp vs ex2,
    summation order).  A flipped decision moves the few affected elements by up to ~1/255 of one pixel's
    contribution, far above 1e-4 of a small element but invisible in the tensor norm.  So: (1) the norm-wise
    relative error must be within `tol`; (2) at least `frac` of the elements must be within `tol` element-wise
    (metric: rel_err); (3) no element may be off by more than `loose`*tol."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if b.size == 0:
        return
    nrm = np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)
    e = elem_err(a, b)
    good = float((e <= tol).mean())
    msg = (f"{what}: norm-wise {nrm:.2e}, elements within {tol:.0e}: {100 * good:.3f}%, worst {e.max():.2e}, "
           f"p99.9 {np.quantile(e, 0.999):.2e}")
    assert nrm <= tol, msg
    assert good >= frac, msg
    assert e.max() <= loose * tol, msg


def oracle_threads():
    """Threads for a live full-size oracle run inside a GPU test (the OpenMP oracle stops scaling past ~32)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(32, n))


def check_image_against_oracle(image, final_T, n_contrib, f, tol=1e-4, min_robust=0.9):
    """The float outputs of the forward against an oracle Forward: image within `tol` on robust pixels (no threshold
    decision within a few ulp of flipping; the oracle flags the others), <= 1e-2 on the flagged ones, n_contrib exact and
    final_T within 1e-3 on robust pixels."""
    ok = f.fragile == 0
    assert ok.mean() > min_robust, f"only {ok.mean():.3f} of the pixels are robust"
    e_img = elem_err(image, f.image).reshape(3, -1)
    assert e_img[:, ok].max() <= tol, f"image: rel err {e_img[:, ok].max():.3e} on a robust pixel"
    assert e_img.max() <= 1e-2, f"image: rel err {e_img.max():.3e} on a fragile pixel"
    nc = np.asarray(n_contrib).reshape(-1)
    assert (nc[ok] == f.n_contrib[ok]).all(), f"n_contrib differs on {(nc[ok] != f.n_contrib[ok]).sum()} robust pixels"
    assert_close(np.asarray(final_T).reshape(-1)[ok], f.final_T[ok], 1e-3, "final_T")
    return float(e_img[:, ok].max())
