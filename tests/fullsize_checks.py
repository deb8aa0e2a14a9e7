"""The body of the full-size GPU test, written against the small surface of divshot_b200.rasterizer.Rasterizer that
it needs (forward / backward / debug_read / stats), so that the CPU suite can run the very same code on an
oracle-backed stand-in (tests/test_props.py) and only the device behaviour itself is left to the GPU run."""
import numpy as np
import torch

import props
from divshot_b200 import _cabi

NAMES = ("means3D", "scales", "quats", "opacities", "sh0", "shN")


def _backward(r, cam, params, sc, dl_host, grad_alloc):
    r.forward(cam, params)
    g = grad_alloc(sc.N, sc.shN.shape[1], r.device)
    g.flat.fill_(float("nan"))  # the kernels must overwrite every element
    r.backward(torch.from_numpy(np.ascontiguousarray(dl_host, dtype=np.float32)).to(r.device), g)
    if r.device.type == "cuda":
        torch.cuda.synchronize()
    return {k: getattr(g, k).cpu().numpy() for k in NAMES}


def run(r, sc, params, gold, linearity, grad_alloc):
    W, H = sc.cameras[0].width, sc.cameras[0].height
    cam = _cabi.make_camera(sc.cameras[0], sc.sh_degree)
    img, radii = r.forward(cam, params)
    radii_h = radii.cpu().numpy()
    rd = r.debug_read
    tt, pl, rg = rd(_cabi.BUF_TILES_TOUCHED), rd(_cabi.BUF_POINT_LIST), rd(_cabi.BUF_RANGES)
    # ---- bit-exact indices vs the oracle's frozen digests
    got = props.index_digests(radii_h, tt, pl, rg)
    assert (got["D"], got["V"]) == (gold["D"], gold["V"]), (got["D"], got["V"], gold["D"], gold["V"])
    for k in ("radii", "tiles_touched", "ranges", "point_list"):
        assert got[k] == gold[k], f"{sc.name}: {k} differs from the oracle at full size"
    # ---- properties
    props.check_binning(pl, rg, rd(_cabi.BUF_DEPTH), radii_h, rd(_cabi.BUF_MEAN2D), tt, W, H)
    nc = rd(_cabi.BUF_N_CONTRIB)
    props.check_compositing(img.cpu().numpy(), rd(_cabi.BUF_FINAL_T), nc, rg, W, H)
    assert (nc > 0).mean() > 0.9, "the synthetic scene covers the image"
    # ---- idempotence; single-pass (deferred-check) binning produces the same lists and image
    img2, _ = r.forward(cam, params)
    assert torch.equal(img, img2) and np.array_equal(rd(_cabi.BUF_POINT_LIST), pl)
    img3, _ = r.forward(cam, params, defer_check=True)
    assert np.array_equal(rd(_cabi.BUF_POINT_LIST), pl) and np.array_equal(rd(_cabi.BUF_RANGES), rg)
    assert torch.equal(img, img3)
    r.stats()  # settles the deferred arena check: raises if the deferred forward overflowed
    # ---- backward
    g1 = _backward(r, cam, params, sc, sc.dL_dpix[0], grad_alloc)
    props.check_gradient_support(g1, radii_h)
    assert all(np.abs(g1[k]).max() > 0 for k in NAMES if g1[k].size)
    if linearity:
        v = np.random.default_rng(5).normal(size=(3, H, W)).astype(np.float32)
        g2 = _backward(r, cam, params, sc, v, grad_alloc)
        g12 = _backward(r, cam, params, sc, 2.0 * sc.dL_dpix[0] - 0.5 * v, grad_alloc)
        props.check_backward_linearity(g1, g2, g12, 2.0, -0.5)
