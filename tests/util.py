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
