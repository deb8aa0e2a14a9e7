"""Test bodies for the trainer's refinement step (divshot_b200/csrc/densify.cu), written once against
a small `Backend` and run twice:
  * tests/test_densify_emul.py — on the CPU, against densify.cu compiled for the host (tests/native/densify_emul.cpp:
    the same kernel bodies as serial loops, the same host orchestration);
  * tests/test_zz_gpu_densify.py — on a B200, against the CUDA kernels in libgstrain.so (marker `gpu`).
Both call the same `dvs_densify_test_*` hooks; only the memory the pointers refer to differs."""
import ctypes as C
import math

import numpy as np

import densify_ref as dr

HOOKS = ("mcmc_refine", "mcmc_noise", "mcmc_regularise", "adc_accumulate", "adc_refine", "adc_reset_opacity")


class Backend:
    """lib: ctypes library exporting dvs_densify_test_*; upload(np) -> handle; ptr(handle) -> int; download(handle) -> np."""

    def __init__(self, lib, upload, ptr, download, sync=lambda: None, exact=False):
        self.lib, self.upload, self.ptr, self.download, self.sync, self.exact = lib, upload, ptr, download, sync, exact
        vp, ll, f, ull = C.c_void_p, C.c_longlong, C.c_float, C.c_ulonglong
        P = C.POINTER
        sig = {
            "mcmc_refine": [vp, vp, vp, P(ll), ll, ll, f, ull, P(ll), vp],
            "mcmc_noise": [vp, ll, f, ull, vp],
            "mcmc_regularise": [vp, vp, ll, f, f, vp],
            "adc_accumulate": [vp, vp, vp, vp, vp, ll, vp],
            "adc_refine": [vp, vp, vp, vp, vp, P(ll), ll, ll, vp, ull, P(ll), vp],
            "adc_reset_opacity": [vp, vp, vp, ll, vp],
        }
        for name, args in sig.items():
            fn = getattr(lib, "dvs_densify_test_" + name)
            fn.argtypes, fn.restype = args, C.c_int


def host_backend(lib):
    return Backend(lib, lambda a: np.ascontiguousarray(a).copy(), lambda h: h.ctypes.data, lambda h: h.copy(), exact=True)


def torch_backend(lib):
    import torch
    return Backend(lib, lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda(), lambda h: h.data_ptr(),
                   lambda h: h.cpu().numpy(), torch.cuda.synchronize)


class _Dev:
    """A model (six arrays) resident on the backend, with the float*[6] table the hooks take."""

    def __init__(self, be, m):
        self.be = be
        self.h = {k: be.upload(m[k]) for k in dr.KEYS}
        self.table = (C.c_void_p * 6)(*[be.ptr(self.h[k]) for k in dr.KEYS])

    def get(self):
        self.be.sync()
        return {k: self.be.download(self.h[k]) for k in dr.KEYS}


def _ok(rc):
    assert rc == 0, f"hook returned cudaError {rc}"


def _mcmc_refine(be, m, m1, m2, N, capacity, cap_max, min_opacity, seed):
    P, M1, M2 = _Dev(be, m), _Dev(be, m1), _Dev(be, m2)
    n, rep = C.c_longlong(N), (C.c_longlong * 6)()
    _ok(be.lib.dvs_densify_test_mcmc_refine(P.table, M1.table, M2.table, C.byref(n), capacity, cap_max, min_opacity, seed, rep, None))
    return P.get(), M1.get(), M2.get(), n.value, list(rep)


def _adc_refine(be, m, m1, m2, accum, denom, N, capacity, cap_max, cfg, seed, revised=False):
    P, M1, M2 = _Dev(be, m), _Dev(be, m1), _Dev(be, m2)
    a, d = be.upload(accum), be.upload(denom)
    c = (C.c_float * 6)(*cfg, 1.0 if revised else 0.0)
    n, rep = C.c_longlong(N), (C.c_longlong * 6)()
    _ok(be.lib.dvs_densify_test_adc_refine(P.table, M1.table, M2.table, be.ptr(a), be.ptr(d), C.byref(n), capacity, cap_max, c, seed, rep, None))
    be.sync()
    return P.get(), M1.get(), M2.get(), be.download(a), be.download(d), n.value, list(rep)


# ------------------------------------------------------------------------------------------------------ MCMC
def case_mcmc_relocation_of_the_dead(be):
    N, cap = 6000, 8000
    before = dr.random_model(N, cap, 1)
    m1b, m2b = dr.moments_like(before, 2.0), dr.moments_like(before, 3.0)
    after, m1a, m2a, N2, rep = _mcmc_refine(be, before, m1b, m2b, N, cap, N, 0.005, 7)  # capMax = N: no growth
    dead = np.flatnonzero(dr.sigmoid(before["opac"][:N]) <= 0.005)
    assert N2 == N and rep[0] == dead.size and rep[1] == dead.size and rep[2] == 0
    cnt = dr.check_relocation(before, after, m1b, m1a, m2b, m2a, N, dead, 0.005, True)
    assert cnt.sum() == dead.size
    for k in dr.KEYS:  # rows beyond N are not touched when nothing is appended
        assert np.array_equal(after[k][N:], before[k][N:]) and np.array_equal(m1a[k][N:], m1b[k][N:])
    # reproducible: the same seed relocates onto the same sources, another seed does not
    again, _, _, _, _ = _mcmc_refine(be, before, m1b, m2b, N, cap, N, 0.005, 7)
    other, _, _, _, _ = _mcmc_refine(be, before, m1b, m2b, N, cap, N, 0.005, 8)
    assert all(np.array_equal(again[k], after[k]) for k in dr.KEYS)
    assert not np.array_equal(other["quats"], after["quats"])


def case_mcmc_growth(be):
    N, cap = 6000, 8000
    before = dr.random_model(N, cap, 1)
    m1b, m2b = dr.moments_like(before, 2.0), dr.moments_like(before, 3.0)
    after, m1a, m2a, N2, rep = _mcmc_refine(be, before, m1b, m2b, N, cap, 10 ** 9, 1e-6, 8)  # nothing is dead
    assert N2 == 6300 and rep[0] == 0 and rep[1] == 0 and rep[2] == 300
    dr.check_relocation(before, after, m1b, m1a, m2b, m2a, N, np.arange(N, N2), 1e-6, False)
    for k in dr.KEYS:
        assert np.array_equal(after[k][N2:], before[k][N2:])
    # growth is bounded by capMax and by the arena capacity; a full arena is left alone
    for cap_max, capacity, expect in ((6100, 8000, 6100), (10 ** 9, 6200, 6200), (5000, 8000, 6000), (6000, 6000, 6000)):
        m = {k: v[:capacity] for k, v in before.items()}
        x, y = ({k: v[:capacity] for k, v in mm.items()} for mm in (m1b, m2b))
        a, xa, ya, n, _ = _mcmc_refine(be, m, x, y, N, capacity, cap_max, 1e-6, 9)
        assert n == expect
        if expect > N:
            dr.check_relocation(m, a, x, xa, y, ya, N, np.arange(N, expect), 1e-6, False)
        else:
            assert all(np.array_equal(a[k], m[k]) for k in dr.KEYS)


def case_mcmc_both_phases_and_edges(be):
    N, cap = 4000, 6000
    before = dr.random_model(N, cap, 5, dead_frac=0.1)
    m1b, m2b = dr.moments_like(before, 2.0), dr.moments_like(before, 3.0)
    after, m1a, m2a, N2, rep = _mcmc_refine(be, before, m1b, m2b, N, cap, 10 ** 9, 0.005, 11)
    dead = np.flatnonzero(dr.sigmoid(before["opac"][:N]) <= 0.005)
    assert N2 == 4200 and rep[:3] == [dead.size, dead.size, 200]
    index = dr._row_index(before["quats"][:N])
    alive_src = 0
    for r in list(dead) + list(range(N, N2)):  # every filled slot is a copy of an input Gaussian, Adam moments zero
        s = index.get(after["quats"][r].tobytes(), -1)
        assert s >= 0 and np.array_equal(after["sh0"][r], before["sh0"][s]) and not m1a["shN"][r].any() and not m2a["opac"][r].any()
        alive_src += dr.sigmoid(before["opac"][s]) > 0.005
    assert alive_src >= dead.size  # the relocation pass only draws live sources
    assert (dr.sigmoid(after["opac"][:N2]) >= 0.005 * (1 - 1e-3)).all(), "nothing is left dead"
    assert all(np.isfinite(after[k][:N2]).all() for k in dr.KEYS)
    # every Gaussian dead: nothing to relocate onto, growth still samples by opacity
    allbad = dr.copy_model(before); allbad["opac"][:N] = -9.0
    a, _, _, n, rep = _mcmc_refine(be, allbad, m1b, m2b, N, cap, N, 0.005, 3)
    assert n == N and rep[0] == N and rep[1] == 0 and all(np.array_equal(a[k], allbad[k]) for k in dr.KEYS)
    # a single Gaussian; an empty model
    one = {k: v[:4].copy() for k, v in before.items()}; one["opac"][0] = 1.0
    a, _, _, n, _ = _mcmc_refine(be, one, {k: v[:4] for k, v in m1b.items()}, {k: v[:4] for k, v in m2b.items()}, 1, 4, 4, 0.005, 3)
    assert n == 1 and np.array_equal(a["means"], one["means"])  # int(1.05 * 1) = 1
    a, _, _, n, _ = _mcmc_refine(be, one, {k: v[:4] for k, v in m1b.items()}, {k: v[:4] for k, v in m2b.items()}, 0, 4, 4, 0.005, 3)
    assert n == 0 and np.array_equal(a["opac"], one["opac"])


def case_mcmc_noise(be, ops):
    from test_densify_ops import _p
    N = 3000
    m = dr.random_model(N, N + 10, 9)
    m["opac"][:N:3] = np.float32(-7.0)  # transparent: these are the ones that move
    P = _Dev(be, m)
    step, seed = 250.0, 99
    _ok(be.lib.dvs_densify_test_mcmc_noise(P.table, N, step, seed, None))
    after = P.get()
    for k in dr.KEYS:
        if k != "means":
            assert np.array_equal(after[k], m[k])
    assert np.array_equal(after["means"][N:], m["means"][N:])
    exp = np.zeros((N, 3), np.float32)
    e, d = np.zeros(4, np.float32), np.zeros(3, np.float32)
    for i in range(N):
        ops.t_normal2(seed, 2 * i, _p(e[:2])); ops.t_normal2(seed, 2 * i + 1, _p(e[2:]))
        ops.t_mcmc_noise(_p(m["scales"][i].copy()), _p(m["quats"][i].copy()), float(m["opac"][i]), _p(e), step, _p(d))
        exp[i] = d
    got = after["means"][:N].astype(np.float64) - m["means"][:N]
    moved = np.abs(exp).max(1) > 1e-6
    assert moved.sum() > N // 4 and np.abs(got[~moved]).max() < 1e-5
    if be.exact:
        assert np.array_equal(after["means"][:N], (m["means"][:N] + exp).astype(np.float32))
    else:  # device expf/logf/sincosf differ from libm by a few ulp; the sum is rounded into the mean
        assert np.allclose(got[moved], exp[moved], rtol=2e-3, atol=2e-6 + 1e-6 * np.abs(m["means"][:N][moved]).max())
    # step 0 and N 0 are no-ops
    Q = _Dev(be, m)
    _ok(be.lib.dvs_densify_test_mcmc_noise(Q.table, N, 0.0, seed, None)); _ok(be.lib.dvs_densify_test_mcmc_noise(Q.table, 0, 1.0, seed, None))
    assert np.array_equal(Q.get()["means"], m["means"])


def case_mcmc_regularise(be):
    N = 5000
    m = dr.random_model(N, N + 7, 4)
    g = {k: np.full_like(v, 0.25) for k, v in m.items()}
    P, G = _Dev(be, m), _Dev(be, g)
    _ok(be.lib.dvs_densify_test_mcmc_regularise(P.table, G.table, N, 0.01, 0.02, None))
    ga = G.get()
    o = dr.sigmoid(m["opac"][:N])
    assert np.allclose(ga["opac"][:N], 0.25 + 0.01 / N * o * (1 - o), rtol=1e-6, atol=0)
    assert np.allclose(ga["scales"][:N], 0.25 + 0.02 / (3 * N) * np.exp(m["scales"][:N].astype(np.float64)), rtol=1e-6, atol=0)
    # the regulariser's gradient is tiny next to 0.25: compare the increments themselves too
    assert np.allclose((ga["opac"][:N].astype(np.float64) - 0.25) * N / 0.01, o * (1 - o), atol=N / 0.01 * 3e-8)
    for k in ("means", "quats", "sh0", "shN"):
        assert np.array_equal(ga[k], g[k])
    assert np.array_equal(ga["opac"][N:], g["opac"][N:]) and np.array_equal(ga["scales"][N:], g["scales"][N:])
    assert all(np.array_equal(v, m[k]) for k, v in P.get().items())


# ------------------------------------------------------------------------------------------------------- ADC
def case_adc_accumulate(be):
    N = 4000
    rng = np.random.default_rng(0)
    g2, ga = rng.normal(size=(N + 5, 2)).astype(np.float32), rng.normal(size=(N + 5, 2)).astype(np.float32)
    radii = rng.integers(-1, 4, N + 5).astype(np.int32)
    acc0, den0 = rng.uniform(0, 1, N + 5).astype(np.float32), rng.integers(0, 9, N + 5).astype(np.float32)
    for use_abs in (False, True):
        hg, ha, hr, acc, den = be.upload(g2), be.upload(ga), be.upload(radii), be.upload(acc0), be.upload(den0)
        _ok(be.lib.dvs_densify_test_adc_accumulate(be.ptr(hg), be.ptr(ha) if use_abs else None, be.ptr(hr), be.ptr(acc), be.ptr(den), N, None))
        be.sync()
        a, d = be.download(acc), be.download(den)
        vis = radii[:N] > 0
        src = (ga if use_abs else g2)[:N].astype(np.float64)
        assert np.allclose(a[:N], acc0[:N] + vis * np.sqrt((src ** 2).sum(1)), rtol=1e-6)
        assert np.array_equal(d[:N], den0[:N] + vis) and np.array_equal(a[:N][~vis], acc0[:N][~vis])
        assert np.array_equal(a[N:], acc0[N:]) and np.array_equal(d[N:], den0[N:])


def _adc_model(N, cap, seed):
    before = dr.random_model(N, cap, seed, dead_frac=0.04)
    before["scales"][:400] = np.log(np.float32(0.2))      # large: split candidates (0.2 > 0.01 * 5)
    before["scales"][400:420] = np.log(np.float32(0.8))   # beyond pruneScale3d * extent = 0.5
    rng = np.random.default_rng(seed + 1)
    denom = rng.integers(0, 20, cap).astype(np.float32)
    accum = (rng.uniform(0, 4e-4, cap) * denom).astype(np.float32)
    return before, accum, denom


def _split_expect(ops, m, seed):
    from test_densify_ops import _p

    def f(i):
        out1, out2 = np.zeros(6, np.float32), np.zeros(6, np.float32)
        for c0, out in ((4 * i, out1), (4 * i + 2, out2)):
            e = np.zeros(4, np.float32)
            ops.t_normal2(seed, c0, _p(e[:2])); ops.t_normal2(seed, c0 + 1, _p(e[2:]))
            ops.t_adc_split_sample(_p(m["means"][i].copy()), _p(m["scales"][i].copy()), _p(m["quats"][i].copy()), _p(e), _p(out))
        return out1[:3].copy(), out2[:3].copy(), out1[3:].copy()
    return f


CFG = (2e-4, 0.01, 5.0, 0.005, 0.1)


def case_adc_refine(be, ops):
    N, cap, seed = 5000, 9000, 1234
    before, accum, denom = _adc_model(N, cap, 2)
    m1b, m2b = dr.moments_like(before, 2.0), dr.moments_like(before, 3.0)
    split = _split_expect(ops, before, seed)
    after, m1a, m2a, acc_a, den_a, N2, rep = _adc_refine(be, before, m1b, m2b, accum, denom, N, cap, 10 ** 9, CFG, seed)
    r = dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N2, cap, 10 ** 9, CFG, split)
    assert r["clones"] > 500 and r["splits"] > 100 and r["pruned"] > 150
    assert rep[3] == r["grown"] and rep[5] == r["pruned"]
    # no room to grow (capMax = N): prune only, holes filled from the tail
    after, m1a, m2a, acc_a, den_a, N3, rep = _adc_refine(be, before, m1b, m2b, accum, denom, N, cap, N, CFG, seed)
    r = dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N3, cap, N, CFG, split)
    assert r["grown"] == 0 and N3 == N - r["pruned"] and rep[3] == 0
    # arena too small for the appended rows although capMax would allow them
    small = {k: v[:N + 50] for k, v in before.items()}
    x, y = ({k: v[:N + 50] for k, v in mm.items()} for mm in (m1b, m2b))
    after, m1a, m2a, acc_a, den_a, N4, _ = _adc_refine(be, small, x, y, accum[:N + 50], denom[:N + 50], N, N + 50, 10 ** 9, CFG, seed)
    r = dr.check_adc_refine(small, after, x, m1a, y, m2a, accum, denom, acc_a, den_a, N, N4, N + 50, 10 ** 9, CFG, split)
    assert r["grown"] == 0


def case_adc_revised_opacity(be, ops):
    """`revisedOpacity`: both Gaussians a clone / split leaves carry 1 - sqrt(1 - o); everything else as without the flag."""
    N, cap, seed = 4000, 8000, 99
    before, accum, denom = _adc_model(N, cap, 4)
    m1b, m2b = dr.moments_like(before, 2.0), dr.moments_like(before, 3.0)
    split = _split_expect(ops, before, seed)
    after, m1a, m2a, acc_a, den_a, N2, rep = _adc_refine(be, before, m1b, m2b, accum, denom, N, cap, 10 ** 9, CFG, seed, revised=True)
    r = dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N2, cap, 10 ** 9, CFG, split, revised=True)
    assert r["clones"] > 300 and r["splits"] > 50
    plain, _, _, _, _, N3, _ = _adc_refine(be, before, m1b, m2b, accum, denom, N, cap, 10 ** 9, CFG, seed)
    assert N3 == N2 and not np.array_equal(plain["opac"][:N2], after["opac"][:N2])
    for k in ("means", "scales", "quats", "sh0", "shN"):
        assert np.array_equal(plain[k][:N2], after[k][:N2]), k  # only the opacities differ
    o = dr.sigmoid(before["opac"][:50]); o2 = dr.sigmoid(dr.revised_opacity_logit(before["opac"][:50]))
    assert np.allclose(1 - (1 - o2) ** 2, o, atol=1e-6)  # the pair composites like the original


def case_adc_edges(be, ops):
    N, cap, seed = 3000, 7000, 77
    before, accum, denom = _adc_model(N, cap, 6)
    m1b, m2b = dr.moments_like(before, 2.0), dr.moments_like(before, 3.0)
    # nothing pruned (thresholds off): clones / splits only, no hole filling
    cfg = (2e-4, 0.01, 5.0, 0.0, 1e9)
    after, m1a, m2a, acc_a, den_a, N2, _ = _adc_refine(be, before, m1b, m2b, accum, denom, N, cap, 10 ** 9, cfg, seed)
    r = dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N2, cap, 10 ** 9, cfg, _split_expect(ops, before, seed))
    assert r["pruned"] == 0 and r["grown"] > 0 and all(np.array_equal(after[k][:N][:5], before[k][:5]) or k in ("means", "scales") for k in dr.KEYS)
    # nothing grows (threshold unreachable): prune only
    cfg = (1e9, 0.01, 5.0, 0.005, 0.1)
    after, m1a, m2a, acc_a, den_a, N2, _ = _adc_refine(be, before, m1b, m2b, accum, denom, N, cap, 10 ** 9, cfg, seed)
    r = dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N2, cap, 10 ** 9, cfg)
    assert r["grown"] == 0 and r["pruned"] > 0
    # everything pruned: the model is empty afterwards
    cfg = (1e9, 0.01, 5.0, 2.0, 0.1)
    _, _, _, _, _, N2, rep = _adc_refine(be, before, m1b, m2b, accum, denom, N, cap, 10 ** 9, cfg, seed)
    assert N2 == 0 and rep[5] == N
    # nothing to do at all: bit-identical model, statistics restarted
    cfg = (1e9, 0.01, 5.0, 0.0, 1e9)
    after, m1a, _, acc_a, den_a, N2, _ = _adc_refine(be, before, m1b, m2b, accum, denom, N, cap, 10 ** 9, cfg, seed)
    assert N2 == N and all(np.array_equal(after[k], before[k]) and np.array_equal(m1a[k], m1b[k]) for k in dr.KEYS)
    assert not acc_a[:N].any() and not den_a[:N].any()
    # pruned Gaussians only at the tail: K = first pruned index, no moves needed
    tail = dr.copy_model(before); tail["opac"][:N] = 1.0; tail["scales"][:N] = -4.0; tail["opac"][N - 100:N] = -9.0
    cfg = (1e9, 0.01, 5.0, 0.005, 0.1)
    after, _, _, _, _, N2, _ = _adc_refine(be, tail, m1b, m2b, accum, denom, N, cap, 10 ** 9, cfg, seed)
    assert N2 == N - 100 and all(np.array_equal(after[k][:N2], tail[k][:N2]) for k in dr.KEYS)


def case_adc_reset_opacity(be):
    N = 2000
    m = dr.random_model(N, N + 9, 8)
    m1, m2 = dr.moments_like(m, 2.0), dr.moments_like(m, 3.0)
    P, M1, M2 = _Dev(be, m), _Dev(be, m1), _Dev(be, m2)
    _ok(be.lib.dvs_densify_test_adc_reset_opacity(P.table, M1.table, M2.table, N, None))
    a, x, y = P.get(), M1.get(), M2.get()
    cap_logit = math.log(0.01 / 0.99)
    assert np.allclose(a["opac"][:N], np.minimum(m["opac"][:N], cap_logit), rtol=1e-6) and (a["opac"][:N] <= cap_logit + 1e-5).all()
    assert not x["opac"][:N].any() and not y["opac"][:N].any()
    assert np.array_equal(a["opac"][N:], m["opac"][N:]) and np.array_equal(x["opac"][N:], m1["opac"][N:])
    for k in dr.KEYS:
        if k != "opac":
            assert np.array_equal(a[k], m[k]) and np.array_equal(x[k], m1[k]) and np.array_equal(y[k], m2[k])


CASES_PLAIN = (case_mcmc_relocation_of_the_dead, case_mcmc_growth, case_mcmc_both_phases_and_edges, case_mcmc_regularise,
               case_adc_accumulate, case_adc_reset_opacity)
CASES_WITH_OPS = (case_mcmc_noise, case_adc_refine, case_adc_edges, case_adc_revised_opacity)
