"""SURVEY.md §8 row F1 — the per-element arithmetic of the trainer's refinement step (divshot_b200/csrc/densify_ops.h,
called by the CUDA kernels of densify.cu), compiled for the host by tests/native/densify_host.cpp and checked against
float64 numpy restatements of the published rules (3DGS-MCMC, arXiv 2404.09591, named at docs/userGuide.md:41; classic
3DGS adaptive density control).  The closed trainer's own implementation is absent from the reference:
no parity claim against it is possible; these tests pin the maths to the papers' formulas."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops():
    src = os.path.join(ROOT, "tests", "native", "densify_host.cpp")
    hdr = os.path.join(ROOT, "divshot_b200", "csrc", "densify_ops.h")
    out = os.path.join(ROOT, "build", "test_densify_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.dirname(hdr), src, "-o", out])
    L = C.CDLL(out)
    L.t_uniform01.restype = C.c_double
    L.t_uniform01.argtypes = [C.c_uint64, C.c_uint64]
    L.t_normal2.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    L.t_sample_cdf.restype = C.c_longlong
    L.t_sample_cdf.argtypes = [C.c_void_p, C.c_longlong, C.c_double]
    L.t_relocation.argtypes = [C.c_float, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    L.t_mcmc_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
    L.t_reg_grad_opacity.restype = C.c_float
    L.t_reg_grad_opacity.argtypes = [C.c_float] * 3
    L.t_reg_grad_scale.restype = C.c_float
    L.t_reg_grad_scale.argtypes = [C.c_float] * 3
    L.t_adc_decide.argtypes = [C.c_float, C.c_float, C.c_void_p] + [C.c_float] * 6
    L.t_adc_split_sample.argtypes = [C.c_void_p] * 5
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _R(q):
    r, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                     [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                     [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])


def test_rng_is_uniform_normal_and_reproducible(ops):
    u = np.array([ops.t_uniform01(42, i) for i in range(20000)])
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.003
    assert ops.t_uniform01(42, 7) == u[7] and ops.t_uniform01(43, 7) != u[7]
    hist, _ = np.histogram(u, 20, (0, 1))
    assert hist.min() > 850 and hist.max() < 1150
    z = np.zeros((10000, 2), np.float32)
    for i in range(10000):
        ops.t_normal2(9, i, _p(z[i]))
    assert np.isfinite(z).all() and abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.03
    assert abs((np.abs(z) < 1).mean() - 0.6827) < 0.01 and abs((np.abs(z) < 2).mean() - 0.9545) < 0.006


def test_cdf_sampling_matches_searchsorted_and_the_weights(ops):
    rng = np.random.default_rng(0)
    w = rng.uniform(0, 1, 500)
    w[rng.integers(0, 500, 120)] = 0.0  # dead entries
    cdf = np.cumsum(w)
    us = rng.uniform(0, cdf[-1], 4000)
    got = np.array([ops.t_sample_cdf(_p(cdf), 500, float(u)) for u in us])
    assert np.array_equal(got, np.searchsorted(cdf, us, side="right"))
    assert (w[got] > 0).all(), "a zero-weight entry must never be drawn"
    assert ops.t_sample_cdf(_p(cdf), 500, 0.0) == int(np.argmax(w > 0))
    assert ops.t_sample_cdf(_p(cdf), 500, float(cdf[-1]) * 2) == 499  # clamped


def _relocation_ref(o, s, n):
    o_new = 1 - (1 - o) ** (1.0 / n)
    denom = sum(math.comb(i - 1, k) * (-1) ** k * o_new ** (k + 1) / math.sqrt(k + 1) for i in range(1, n + 1) for k in range(i))
    return o_new, np.asarray(s, np.float64) * o / denom


def test_relocation_rule_matches_the_paper_formula(ops):
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 5, 8, 13, 21):
        for _ in range(20):
            o = float(rng.uniform(0.02, 0.97))
            s = rng.uniform(0.001, 0.5, 3).astype(np.float32)
            out = np.zeros(4, np.float32)
            ops.t_relocation(o, _p(s), n, 0.005, _p(out))
            o_ref, s_ref = _relocation_ref(np.float32(o), s, n)
            o_ref = min(max(o_ref, 0.005), 1 - 2.0 ** -24)
            assert abs(1 / (1 + math.exp(-out[0])) - o_ref) < 2e-5 * max(o_ref, 0.05), (n, o)
            assert np.allclose(np.exp(out[1:].astype(np.float64)), s_ref, rtol=5e-4 if n <= 8 else 5e-3), (n, o)
    # n = 1 is the identity
    s = np.array([0.1, 0.2, 0.3], np.float32); out = np.zeros(4, np.float32)
    ops.t_relocation(0.4, _p(s), 1, 0.005, _p(out))
    assert abs(1 / (1 + math.exp(-out[0])) - 0.4) < 1e-6 and np.allclose(np.exp(out[1:]), s, rtol=1e-5)
    # the n copies together are as opaque as the original: 1 - (1 - o')^n = o
    ops.t_relocation(0.6, _p(s), 4, 0.005, _p(out))
    o4 = 1 / (1 + math.exp(-out[0]))
    assert abs(1 - (1 - o4) ** 4 - 0.6) < 1e-5
    # ratios beyond the table are clamped to 51 copies, tiny opacities to min_opacity
    a = np.zeros(4, np.float32); b = np.zeros(4, np.float32)
    ops.t_relocation(0.3, _p(s), 51, 0.005, _p(a)); ops.t_relocation(0.3, _p(s), 500, 0.005, _p(b))
    assert np.array_equal(a, b)
    assert abs(1 / (1 + math.exp(-a[0])) - max(1 - 0.7 ** (1 / 51), 0.005)) < 1e-6


def test_exploration_noise_is_covariance_shaped_and_gated_by_opacity(ops):
    rng = np.random.default_rng(2)
    for _ in range(50):
        ls = rng.normal(-3, 1, 3).astype(np.float32)
        q = rng.normal(size=4).astype(np.float32)
        eps = rng.normal(size=3).astype(np.float32)
        logit = float(rng.normal(-5.5, 1.5))
        d = np.zeros(3, np.float32)
        ops.t_mcmc_noise(_p(ls), _p(q), logit, _p(eps), 0.37, _p(d))
        R = _R(q.astype(np.float64))
        sigma = R @ np.diag(np.exp(2 * ls.astype(np.float64))) @ R.T
        o = 1 / (1 + math.exp(-logit))
        gate = 1 / (1 + math.exp(-100 * ((1 - o) - 0.995)))
        ref = sigma @ eps.astype(np.float64) * gate * 0.37
        assert np.allclose(d, ref, rtol=2e-4, atol=1e-7 * np.abs(ref).max() + 1e-30)
    # an opaque Gaussian does not move (gate ~ e^-49 at opacity 0.5)
    ops.t_mcmc_noise(_p(ls), _p(q), 0.0, _p(eps), 1e3, _p(d))
    assert np.abs(d).max() < 1e-15


def test_regulariser_gradients_match_autograd(ops):
    import torch
    x = torch.tensor([-3.0, -0.5, 0.0, 2.0], dtype=torch.float64, requires_grad=True)
    ls = torch.tensor([-5.0, -2.0, 0.3], dtype=torch.float64, requires_grad=True)
    (0.01 * torch.sigmoid(x).mean() + 0.02 * torch.exp(ls).mean()).backward()
    for v, g in zip(x.tolist(), x.grad.tolist()):
        assert abs(ops.t_reg_grad_opacity(v, 0.01, 1 / 4) - g) < 1e-8
    for v, g in zip(ls.tolist(), ls.grad.tolist()):
        assert abs(ops.t_reg_grad_scale(v, 0.02, 1 / 3) - g) < 1e-7 * max(1, abs(g))


def test_adc_decisions_and_split_samples(ops):
    KEEP, CLONE, SPLIT, PRUNE = 0, 1, 2, 4
    small, big, huge = np.log(np.float32([0.005] * 3)), np.log(np.float32([0.005, 0.2, 0.01])), np.log(np.float32([3.0, 0.1, 0.1]))
    dec = lambda acc, den, ls, logit: ops.t_adc_decide(acc, den, _p(np.ascontiguousarray(ls, np.float32)), logit, 2e-4, 0.01, 5.0, 0.005, 0.1)
    assert dec(1e-3, 10, small, 0.0) == KEEP              # mean gradient 1e-4 < threshold
    assert dec(3e-3, 10, small, 0.0) == CLONE             # 3e-4 >= 2e-4, max scale 0.005 <= 0.05
    assert dec(3e-3, 10, big, 0.0) == SPLIT               # max scale 0.2 > 0.05
    assert dec(3e-3, 0, big, 0.0) == KEEP                 # never visible
    assert dec(0.0, 10, small, -6.0) == PRUNE             # opacity 0.0025 < 0.005
    assert dec(3e-3, 10, huge, 0.0) == (SPLIT | PRUNE)    # max scale 3 > 0.1 * 5
    rng = np.random.default_rng(3)
    mean, ls, q, eps = rng.normal(size=3).astype(np.float32), rng.normal(-2, 0.5, 3).astype(np.float32), rng.normal(size=4).astype(np.float32), rng.normal(size=3).astype(np.float32)
    out = np.zeros(6, np.float32)
    ops.t_adc_split_sample(_p(mean), _p(ls), _p(q), _p(eps), _p(out))
    ref = mean + _R(q.astype(np.float64)) @ (np.exp(ls.astype(np.float64)) * eps)
    assert np.allclose(out[:3], ref, rtol=1e-5, atol=1e-6) and np.allclose(np.exp(out[3:]), np.exp(ls) / 1.6, rtol=1e-6)
