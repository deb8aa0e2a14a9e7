"""The refinement-step checkers of tests/densify_ref.py must accept a faithful numpy emulation of the step and reject
corrupted results — proven here on the CPU so that the GPU tests (tests/test_zz_gpu_densify.py) can aim them
at divshot_b200/csrc/densify.cu."""
import ctypes as C

import numpy as np
import pytest

import densify_ref as dr
from test_densify_ops import _p, ops  # noqa: F401  (host build of densify_ops.h)


def _split_samples(ops, m, seed):  # noqa: F811
    def f(i):
        out1, out2 = np.zeros(6, np.float32), np.zeros(6, np.float32)
        for c0, out in ((4 * i, out1), (4 * i + 2, out2)):
            e = np.zeros(4, np.float32)
            ops.t_normal2(seed, c0, _p(e[:2])); ops.t_normal2(seed, c0 + 1, _p(e[2:]))
            ops.t_adc_split_sample(_p(m["means"][i].copy()), _p(m["scales"][i].copy()), _p(m["quats"][i].copy()), _p(e), _p(out))
        return out1[:3].copy(), out2[:3].copy(), out1[3:].copy()
    return f


def test_relocation_checker_accepts_the_emulation_and_rejects_corruption():
    N, cap = 6000, 8000
    before = dr.random_model(N, cap, 1)
    m1b, m2b = dr.moments_like(before, 2.0), dr.moments_like(before, 3.0)
    # phase 1 alone (no room to grow)
    after, m1a, m2a = dr.copy_model(before), dr.copy_model(m1b), dr.copy_model(m2b)
    assert dr.emulate_mcmc_refine(after, m1a, m2a, N, cap, N, 0.005, 7) == N
    dead = np.flatnonzero(dr.sigmoid(before["opac"][:N]) <= 0.005)
    assert 300 <= dead.size < 360  # 5 % planted + the natural tail of the opacity distribution
    cnt = dr.check_relocation(before, after, m1b, m1a, m2b, m2a, N, dead, 0.005, True)
    assert cnt.sum() == dead.size and cnt.max() >= 2
    bad = dr.copy_model(after); bad["opac"][np.flatnonzero(cnt)[0]] += 0.01
    with pytest.raises(AssertionError):
        dr.check_relocation(before, bad, m1b, m1a, m2b, m2a, N, dead, 0.005, True)
    bad = dr.copy_model(after); bad["quats"][dead[3]] = before["quats"][dead[4]]  # a copy of a dead Gaussian
    with pytest.raises(AssertionError):
        dr.check_relocation(before, bad, m1b, m1a, m2b, m2a, N, dead, 0.005, True)
    badm = dr.copy_model(m1a); badm["shN"][dead[0]] = 1
    with pytest.raises(AssertionError):
        dr.check_relocation(before, after, m1b, badm, m2b, m2a, N, dead, 0.005, True)
    bad = dr.copy_model(after); bad["means"][int(np.argmax(cnt == 0))] += 1  # an untouched Gaussian moved
    with pytest.raises(AssertionError):
        dr.check_relocation(before, bad, m1b, m1a, m2b, m2a, N, dead, 0.005, True)
    # phase 2 alone (nothing dead at min_opacity = 1e-6)
    after, m1a, m2a = dr.copy_model(before), dr.copy_model(m1b), dr.copy_model(m2b)
    N2 = dr.emulate_mcmc_refine(after, m1a, m2a, N, cap, 10 ** 9, 1e-6, 8)
    assert N2 == 6300
    dr.check_relocation(before, after, m1b, m1a, m2b, m2a, N, np.arange(N, N2), 1e-6, False)
    # growth is bounded by capMax and by the arena capacity
    for cap_max, capacity, expect in ((6100, 8000, 6100), (10 ** 9, 6200, 6200), (5000, 8000, 6000)):
        a, x, y = dr.copy_model(before), dr.copy_model(m1b), dr.copy_model(m2b)
        assert dr.emulate_mcmc_refine(a, x, y, N, capacity, cap_max, 1e-6, 9) == expect


def test_adc_checker_accepts_the_emulation_and_rejects_corruption(ops):  # noqa: F811
    N, cap, seed = 5000, 9000, 1234
    before = dr.random_model(N, cap, 2, dead_frac=0.04)
    before["scales"][:400] = np.log(np.float32(0.2))      # large: split candidates (0.2 > 0.01 * 5)
    before["scales"][400:420] = np.log(np.float32(0.8))   # beyond pruneScale3d * extent = 0.5
    rng = np.random.default_rng(3)
    denom = rng.integers(0, 20, cap).astype(np.float32)
    accum = (rng.uniform(0, 4e-4, cap) * denom).astype(np.float32)
    cfg = (2e-4, 0.01, 5.0, 0.005, 0.1)
    m1b, m2b = dr.moments_like(before, 2.0), dr.moments_like(before, 3.0)
    after, m1a, m2a, acc_a, den_a = dr.copy_model(before), dr.copy_model(m1b), dr.copy_model(m2b), accum.copy(), denom.copy()
    split = _split_samples(ops, before, seed)
    N2 = dr.emulate_adc_refine(after, m1a, m2a, acc_a, den_a, N, cap, 10 ** 9, cfg, split)
    rep = dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N2, cap, 10 ** 9, cfg, split)
    assert rep["clones"] > 500 and rep["splits"] > 100 and rep["pruned"] > 150 and N2 == N + rep["grown"] - rep["pruned"]
    bad = dr.copy_model(after); bad["means"][N2 - 1] += 1
    with pytest.raises(AssertionError):
        dr.check_adc_refine(before, bad, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N2, cap, 10 ** 9, cfg, split)
    with pytest.raises(AssertionError):
        dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N2 - 1, cap, 10 ** 9, cfg, split)
    bad_acc = acc_a.copy(); bad_acc[5] = 1
    with pytest.raises(AssertionError):
        dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, bad_acc, den_a, N, N2, cap, 10 ** 9, cfg, split)
    # no room to grow: prune only
    after, m1a, m2a, acc_a, den_a = dr.copy_model(before), dr.copy_model(m1b), dr.copy_model(m2b), accum.copy(), denom.copy()
    N3 = dr.emulate_adc_refine(after, m1a, m2a, acc_a, den_a, N, cap, N, cfg, split)
    rep = dr.check_adc_refine(before, after, m1b, m1a, m2b, m2a, accum, denom, acc_a, den_a, N, N3, cap, N, cfg, split)
    assert rep["grown"] == 0 and N3 == N - rep["pruned"]
