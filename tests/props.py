"""Size-independent properties of the rasterizer's outputs (numpy only; the inputs may come from the CUDA path or from
the oracle — the CPU tests run every checker on oracle outputs, including corrupted ones, before the GPU tests trust
them at BASELINE.json's full sizes where a live oracle comparison of every float is too slow).

Every checker raises AssertionError with a message naming the first violation.
Conventions: point_list[D] sorted Gaussian ids, ranges[T,2] = [first, last+1) per tile (untouched tiles (0,0)),
depth[N] fp32 view depth, radii[N] int32 (0 = invisible), mean2D[N,2] fp32 pixel centres, tiles_touched[N] uint32.
"""
import hashlib

import numpy as np

TILE = 16


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def index_digests(radii, tiles_touched, point_list, ranges) -> dict:
    """sha256 of the four integer outputs that must be bit-exact (north_star: 'bit-exact on tile/sort indices')."""
    return {"radii": sha(radii.astype(np.int32)), "tiles_touched": sha(tiles_touched.astype(np.uint32)),
            "point_list": sha(point_list.astype(np.uint32)), "ranges": sha(ranges.astype(np.uint32)),
            "D": int(point_list.shape[0]), "V": int((radii > 0).sum())}


def tile_rects(mean2D, radii, width, height):
    """Tile rectangle of every Gaussian, recomputed from (mean2D, radius) with the literal fp32 arithmetic of
    SURVEY.md Appendix B.1 step 7: min = clamp((int)((mu - r)/16), 0, G), max = clamp((int)((mu + r + 15)/16), 0, G)."""
    gx, gy = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    r = radii.astype(np.float32)
    f16 = np.float32(16.0)

    def lo(c):
        return ((mean2D[:, c] - r) / f16).astype(np.int32)  # astype truncates toward zero like the C cast

    def hi(c):
        return ((mean2D[:, c] + r + np.float32(15.0)) / f16).astype(np.int32)

    rect = np.stack([np.clip(lo(0), 0, gx), np.clip(lo(1), 0, gy), np.clip(hi(0), 0, gx), np.clip(hi(1), 0, gy)], 1)
    rect[radii <= 0] = 0
    return rect, gx, gy


def check_binning(point_list, ranges, depth, radii, mean2D, tiles_touched, width, height):
    """(1) the tile ranges partition [0, D) in tile order; (2) every tile's list is sorted by (depth bits, id);
    (3) per-tile list lengths equal the number of Gaussians whose tile rectangle covers the tile (a 2-D histogram
    rebuilt from mean2D/radii — the 'checksum of checksums' of the duplication stage); (4) tiles_touched = rect area,
    sum = D; (5) every id in a tile's list really covers that tile."""
    D = point_list.shape[0]
    rect, gx, gy = tile_rects(mean2D, radii, width, height)
    area = ((rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1])).astype(np.int64)
    vis = radii > 0
    assert np.array_equal(area[vis], tiles_touched[vis].astype(np.int64)), "tiles_touched != tile-rectangle area"
    assert not tiles_touched[~vis].any(), "invisible Gaussian with tiles_touched != 0"
    assert int(area.sum()) == D, f"sum(tiles_touched) = {int(area.sum())} but the list holds {D} entries"
    ranges = ranges.astype(np.int64)
    T = gx * gy
    assert ranges.shape == (T, 2)
    length = ranges[:, 1] - ranges[:, 0]
    assert (length >= 0).all(), "negative tile range"
    touched = length > 0
    starts = ranges[touched, 0]
    if starts.size:
        assert np.array_equal(starts, np.concatenate([[0], np.cumsum(length[touched])[:-1]])), "ranges do not partition [0, D) in tile order"
    assert int(length.sum()) == D
    assert not ranges[~touched].any(), "untouched tile with a non-zero range"
    # expected per-tile counts from the rectangles: 2-D difference array
    diff = np.zeros((gy + 1, gx + 1), np.int64)
    v = rect[vis]
    np.add.at(diff, (v[:, 1], v[:, 0]), 1)
    np.add.at(diff, (v[:, 1], v[:, 2]), -1)
    np.add.at(diff, (v[:, 3], v[:, 0]), -1)
    np.add.at(diff, (v[:, 3], v[:, 2]), 1)
    expect = diff.cumsum(0).cumsum(1)[:gy, :gx].reshape(-1)
    assert np.array_equal(expect, length), f"per-tile list lengths differ from the rectangle histogram at tile {int(np.argmax(expect != length))}"
    if D == 0:
        return
    ids = point_list.astype(np.int64)
    assert ids.min() >= 0 and ids.max() < radii.shape[0], "id out of range"
    tile_of = np.repeat(np.arange(T, dtype=np.int64), length)
    key = (depth.view(np.uint32)[ids].astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
    same = tile_of[1:] == tile_of[:-1]
    bad = same & (key[1:] <= key[:-1])  # strictly increasing: an id appears at most once per tile
    assert not bad.any(), f"tile list not sorted by (depth bits, id) at entry {int(np.argmax(bad)) + 1}"
    tx, ty = tile_of % gx, tile_of // gx
    rr = rect[ids]
    inside = (tx >= rr[:, 0]) & (tx < rr[:, 2]) & (ty >= rr[:, 1]) & (ty < rr[:, 3])
    assert inside.all(), f"entry {int(np.argmin(inside))} lists a Gaussian whose rectangle misses the tile"


def check_compositing(image, final_T, n_contrib, ranges, width, height, bg=(0, 0, 0)):
    """final_T in [1e-4 * (1 - 0.99), 1]: transmittance only ever shrinks by factors (1 - alpha) >= 0.01 and a pixel
    stops before it would fall under 1e-4; n_contrib <= its tile's list length; n_contrib == 0 <=> the pixel shows
    exactly the background with T = 1; colours are finite and non-negative for a non-negative background."""
    gx = (width + TILE - 1) // TILE
    P = width * height
    assert image.shape == (3, height, width) and final_T.shape[0] == P and n_contrib.shape[0] == P
    assert np.isfinite(image).all() and np.isfinite(final_T).all()
    assert (final_T <= 1.0).all() and (final_T >= 1e-4).all(), "final_T outside [1e-4, 1]"
    ys, xs = np.divmod(np.arange(P), width)
    tile = (ys // TILE) * gx + xs // TILE
    length = (ranges[:, 1].astype(np.int64) - ranges[:, 0].astype(np.int64))[tile]
    assert (n_contrib.astype(np.int64) <= length).all(), "n_contrib beyond the tile list"
    empty = n_contrib == 0
    assert (final_T[empty] == 1.0).all(), "pixel without contributors must keep T = 1"
    flat = image.reshape(3, P)
    for c in range(3):
        assert (flat[c, empty] == np.float32(bg[c])).all(), "pixel without contributors must show the background"
    if min(bg) >= 0:
        assert (flat >= 0).all(), "negative colour"


def check_gradient_support(grads: dict, radii):
    """Invisible Gaussians receive exactly zero gradient; everything is finite."""
    inv = radii <= 0
    for k, g in grads.items():
        assert np.isfinite(g).all(), f"{k}: non-finite gradient"
        assert not g[inv].any(), f"{k}: invisible Gaussian with a non-zero gradient"


def norm_err(a, b) -> float:
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def check_backward_linearity(g1: dict, g2: dict, g12: dict, a: float, b: float, tol=1e-4):
    """The backward pass is linear in dL/dpixel: grads(a*u + b*v) = a*grads(u) + b*grads(v) (norm-wise, per tensor;
    summation order differs between runs, so not bit-exact)."""
    for k in g12:
        e = norm_err(g12[k], a * g1[k].astype(np.float64) + b * g2[k].astype(np.float64))
        assert e <= tol, f"{k}: backward not linear in dL/dpix (norm-wise error {e:.2e})"
