"""Checkers for the trainer's refinement step (divshot_b200/csrc/densify.cu) plus a numpy emulation of
its semantics.  The emulation exists so that the CPU suite can prove the CHECKERS right (they must accept the
emulation's output and reject corrupted copies, tests/test_densify_ref.py) before the GPU tests
(tests/test_zz_gpu_densify.py) point them at the device code.  A model is a dict of float32 arrays
{means[cap,3], scales[cap,3], quats[cap,4], opac[cap], sh0[cap,3], shN[cap,45]} of which the first N rows are live."""
import math

import numpy as np

KEYS = ("means", "scales", "quats", "opac", "sh0", "shN")
WIDTH = {"means": 3, "scales": 3, "quats": 4, "opac": 0, "sh0": 3, "shN": 45}


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x, np.float64)))


def random_model(N, cap, seed, dead_frac=0.05):
    rng = np.random.default_rng(seed)
    m = {k: (rng.normal(size=(cap, WIDTH[k])) if WIDTH[k] else rng.normal(size=cap)).astype(np.float32) for k in KEYS}
    m["scales"] = rng.normal(-3.5, 0.7, (cap, 3)).astype(np.float32)
    m["opac"] = rng.normal(0.0, 2.0, cap).astype(np.float32)
    dead = rng.choice(N, int(dead_frac * N), replace=False)
    m["opac"][dead] = rng.uniform(-9.0, -5.4, dead.size).astype(np.float32)  # sigmoid <= 0.0045
    return m


def moments_like(m, scale):
    """Adam-moment arenas with a recognisable, non-zero content derived from the parameters."""
    return {k: (scale * m[k] + np.float32(1.0)).astype(np.float32) for k in KEYS}


def copy_model(m):
    return {k: v.copy() for k, v in m.items()}


def relocation(o, s, n, min_opacity):
    """Eq. 9 of 3DGS-MCMC in float64: opacity/scale of each of n copies -> (logit, log-scale[3])."""
    n = int(min(max(n, 1), 51))
    o_new = 1.0 - (1.0 - o) ** (1.0 / n)
    denom = sum(math.comb(i - 1, k) * (-1) ** k * o_new ** (k + 1) / math.sqrt(k + 1) for i in range(1, n + 1) for k in range(i))
    oc = min(max(o_new, min_opacity), 1.0 - 2.0 ** -24)
    return math.log(oc / (1.0 - oc)), np.log(np.asarray(s, np.float64) * o / denom)


def _apply_relocation(m, m1, m2, src, dst, min_opacity):
    """Sources `src` (with multiplicity) are copied onto slots `dst`; sources and copies get the relocation values."""
    cnt = np.bincount(src, minlength=m["opac"].shape[0])
    new = {}
    for s in np.flatnonzero(cnt):
        new[s] = relocation(float(sigmoid(m["opac"][s])), np.exp(m["scales"][s].astype(np.float64)), cnt[s] + 1, min_opacity)
    for s, d in zip(src, dst):
        for k in KEYS:
            m[k][d] = m[k][s]
        m["opac"][d] = np.float32(new[s][0]); m["scales"][d] = new[s][1].astype(np.float32)
        for mm in (m1, m2):
            for k in KEYS:
                mm[k][d] = 0
    for s in new:
        m["opac"][s] = np.float32(new[s][0]); m["scales"][s] = new[s][1].astype(np.float32)
        for mm in (m1, m2):
            for k in KEYS:
                mm[k][s] = 0


def emulate_mcmc_refine(m, m1, m2, N, capacity, cap_max, min_opacity, seed):
    rng = np.random.default_rng(seed)
    o = sigmoid(m["opac"][:N])
    dead = np.flatnonzero(o <= min_opacity)
    if 0 < dead.size < N:
        w = np.where(o <= min_opacity, 0.0, o)
        _apply_relocation(m, m1, m2, rng.choice(N, dead.size, p=w / w.sum()), dead, min_opacity)
    target = min(cap_max, capacity, int(1.05 * N))
    n_new = max(0, target - N)
    if n_new:
        o = sigmoid(m["opac"][:N])
        _apply_relocation(m, m1, m2, rng.choice(N, n_new, p=o / o.sum()), np.arange(N, N + n_new), min_opacity)
    return N + n_new


def _row_index(rows):
    return {r.tobytes(): i for i, r in enumerate(rows)}


def check_relocation(before, after, m1b, m1a, m2b, m2a, N, dst, min_opacity, sources_must_be_alive, rtol=2e-4):
    """One relocation pass judged from the state before and after: every slot of `dst` (the dead slots, or the appended
    range) is a copy of a source in [0, N) \\ dst; a source sampled c times and its c copies all hold
    relocation(o, s, c + 1); their Adam moments are zero; every other live Gaussian is bit-identical, moments included.
    `quats` rows identify a Gaussian (unique in the test model, never changed by this step)."""
    dst = np.asarray(dst, int)
    index = _row_index(before["quats"][:N])
    assert len(index) == N, "test model must have unique quaternion rows"
    o_b = sigmoid(before["opac"][:N])
    is_dst = np.zeros(max(N, int(dst.max()) + 1 if dst.size else N), bool)
    is_dst[dst] = True
    src = np.empty(dst.size, int)
    for j, d in enumerate(dst):
        s = index.get(after["quats"][d].tobytes(), -1)
        assert s >= 0 and not is_dst[s], f"slot {d} is not a copy of a valid source"
        if sources_must_be_alive:
            assert o_b[s] > min_opacity, f"slot {d} copies the dead Gaussian {s}"
        src[j] = s
    cnt = np.bincount(src, minlength=N)
    for s in np.flatnonzero(cnt):
        logit, ls = relocation(float(o_b[s]), np.exp(before["scales"][s].astype(np.float64)), cnt[s] + 1, min_opacity)
        rows = [s] + [int(d) for d in dst[src == s]]
        for r in rows:
            assert abs(float(after["opac"][r]) - logit) <= rtol * max(1.0, abs(logit)), f"opacity of row {r} (source {s}, {cnt[s]} copies)"
            assert np.allclose(after["scales"][r], ls, rtol=0, atol=rtol * 5), f"scale of row {r} (source {s})"
            for k in ("means", "sh0", "shN"):
                assert np.array_equal(after[k][r], before[k][s]), (k, r)
            for mm in (m1a, m2a):
                assert all(not np.any(mm[k][r]) for k in KEYS), f"Adam moments of row {r} not zeroed"
    untouched = np.ones(N, bool)
    untouched[np.flatnonzero(cnt)] = False
    untouched[dst[dst < N]] = False
    for k in KEYS:
        assert np.array_equal(after[k][:N][untouched], before[k][:N][untouched]), f"untouched {k} changed"
        assert np.array_equal(m1a[k][:N][untouched], m1b[k][:N][untouched]), f"untouched m1.{k} changed"
        assert np.array_equal(m2a[k][:N][untouched], m2b[k][:N][untouched]), f"untouched m2.{k} changed"
    if dst.size >= 200:  # sources are drawn with probability ~ opacity: their mean opacity is E[o^2]/E[o] > E[o]
        pool = o_b[(o_b > min_opacity) & ~is_dst[:N]] if sources_must_be_alive else o_b[~is_dst[:N]]
        expect = (pool ** 2).sum() / pool.sum()
        assert abs(o_b[src].mean() - expect) < 0.15 * expect, (o_b[src].mean(), expect, pool.mean())
    return cnt


# ---------------------------------------------------------------------------------------------------- ADC
ADC_CLONE, ADC_SPLIT, ADC_PRUNE = 1, 2, 4


def adc_actions(m, accum, denom, N, cfg):
    thr, pd, extent, po, ps = cfg
    g = np.where(denom[:N] > 0, accum[:N] / np.maximum(denom[:N], 1e-30), 0.0)
    smax = np.exp(m["scales"][:N].max(1).astype(np.float64))
    act = np.zeros(N, np.uint8)
    grow = g >= thr
    act[grow & (smax <= pd * extent)] = ADC_CLONE
    act[grow & (smax > pd * extent)] = ADC_SPLIT
    act[(sigmoid(m["opac"][:N]) < po) | (smax > ps * extent)] |= ADC_PRUNE
    return act


def revised_opacity_logit(logit):
    """`revisedOpacity` (arXiv 2404.06109 eq. 9): each of the two Gaussians a clone / split leaves gets 1 - sqrt(1 - o)."""
    o2 = 1.0 - np.sqrt(1.0 - sigmoid(logit))
    return np.log(o2 / (1.0 - o2))


def check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum_b, denom_b, accum_a, denom_a, N, N_after, capacity,
                     cap_max, cfg, split_expect=None, revised=False):
    """split_expect(i) -> (mean1, mean2, log_scale) of the two samples of split Gaussian i (from the host build of the
    same per-element code), or None to only check the scale."""
    act = adc_actions(before, accum_b, denom_b, N, cfg)
    pruned = (act & ADC_PRUNE) != 0
    grow = ~pruned & ((act & (ADC_CLONE | ADC_SPLIT)) != 0)
    n_grow, n_pruned = int(grow.sum()), int(pruned.sum())
    if N - n_pruned + n_grow > min(cap_max, capacity) or N + n_grow > capacity:
        n_grow, grow = 0, np.zeros(N, bool)
    assert N_after == N + n_grow - n_pruned, f"N after = {N_after}, expected {N + n_grow - n_pruned}"
    index = _row_index(before["quats"][:N])
    assert len(index) == N
    seen = np.zeros(N, int)
    rows_of = [[] for _ in range(N)]
    for r in range(N_after):
        i = index.get(after["quats"][r].tobytes(), -1)
        assert i >= 0, f"row {r} is not derived from any input Gaussian"
        seen[i] += 1
        rows_of[i].append(r)
    assert (seen[pruned] == 0).all(), "a pruned Gaussian survived"
    assert (seen[grow] == 2).all(), "a cloned/split Gaussian must appear twice"
    assert (seen[~pruned & ~grow] == 1).all(), "an untouched Gaussian must appear exactly once"
    for i in range(N):
        if pruned[i]:
            continue
        for r in rows_of[i]:
            for k in ("sh0", "shN"):
                assert np.array_equal(after[k][r], before[k][i]), (k, i)
            if revised and grow[i]:
                assert abs(float(after["opac"][r]) - float(revised_opacity_logit(before["opac"][i]))) <= 2e-4 * max(1.0, abs(float(before["opac"][i]))), ("revised opacity", i)
            else:
                assert np.array_equal(after["opac"][r], before["opac"][i]), ("opac", i)
        if grow[i] and act[i] & ADC_SPLIT:
            for r in rows_of[i]:
                assert np.allclose(after["scales"][r], before["scales"][i] - math.log(1.6), atol=1e-5), "split scale"
                for mm in (m1a, m2a):
                    assert all(not np.any(mm[k][r]) for k in KEYS), "Adam moments of a split sample not zeroed"
            got = sorted([tuple(np.round(after["means"][r].astype(np.float64), 4)) for r in rows_of[i]])
            assert got[0] != got[1], "the two split samples must differ"
            if split_expect is not None:
                m_1, m_2, _ = split_expect(i)
                exp = sorted([tuple(np.round(np.asarray(m_1, np.float64), 4)), tuple(np.round(np.asarray(m_2, np.float64), 4))])
                assert np.allclose(got, exp, atol=2e-4), f"split samples of Gaussian {i}"
        else:
            for r in rows_of[i]:
                assert np.array_equal(after["means"][r], before["means"][i]) and np.array_equal(after["scales"][r], before["scales"][i])
            # exactly one row keeps the source's Adam moments; a clone's moments are zero
            kept = [r for r in rows_of[i] if all(np.array_equal(m1a[k][r], m1b[k][i]) and np.array_equal(m2a[k][r], m2b[k][i]) for k in KEYS)]
            assert len(kept) == 1, f"Gaussian {i}: {len(kept)} rows carry its Adam moments"
            for r in rows_of[i]:
                if r not in kept:
                    assert all(not np.any(m1a[k][r]) and not np.any(m2a[k][r]) for k in KEYS), "clone moments not zeroed"
    assert not accum_a[:N_after].any() and not denom_a[:N_after].any(), "statistics must restart after a refinement"
    return dict(grown=n_grow, pruned=n_pruned, clones=int((grow & ((act & ADC_CLONE) != 0)).sum()),
                splits=int((grow & ((act & ADC_SPLIT) != 0)).sum()))


def emulate_adc_refine(m, m1, m2, accum, denom, N, capacity, cap_max, cfg, split_samples, revised=False):
    """split_samples(i) -> (mean1, mean2, log_scale).  Same layout policy as the device code: appended elements at
    [N, N+n_grow), then holes below K filled from the tail in ascending order."""
    act = adc_actions(m, accum, denom, N, cfg)
    pruned = (act & ADC_PRUNE) != 0
    grow = np.flatnonzero(~pruned & ((act & (ADC_CLONE | ADC_SPLIT)) != 0))
    n_pruned = int(pruned.sum())
    if N - n_pruned + grow.size > min(cap_max, capacity) or N + grow.size > capacity:
        grow = grow[:0]
    for j, s in enumerate(grow):
        d = N + j
        for k in KEYS:
            m[k][d] = m[k][s]; m1[k][d] = 0; m2[k][d] = 0
        if revised:
            m["opac"][d] = m["opac"][s] = np.float32(revised_opacity_logit(m["opac"][s]))
        if act[s] & ADC_SPLIT:
            a, b, ls = split_samples(s)
            m["means"][d] = b; m["scales"][d] = ls
            m["means"][s] = a; m["scales"][s] = ls
            for k in KEYS:
                m1[k][s] = 0; m2[k][s] = 0
    total = N + grow.size
    K = total - n_pruned
    pr = np.zeros(total, bool); pr[:N] = pruned
    holes = np.flatnonzero(pr[:K]); movers = K + np.flatnonzero(~pr[K:])
    for d, s in zip(holes, movers):
        for k in KEYS:
            m[k][d] = m[k][s]; m1[k][d] = m1[k][s]; m2[k][d] = m2[k][s]
    accum[:total] = 0; denom[:total] = 0
    return K
