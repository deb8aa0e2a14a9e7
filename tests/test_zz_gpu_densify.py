"""The refinement step of divshot_b200/csrc/densify.cu on a B200 (marker `gpu`; green on B200 since the round-1 driver run),
through the dvs_densify_test_* hooks of libgstrain.so.  The same test bodies (tests/densify_cases.py) pass on the CPU
against the host build of the same source (tests/test_densify_emul.py)."""
import ctypes as C
import os

import pytest

import densify_cases as dc
from test_densify_ops import ops  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    import torch
    assert torch.cuda.is_available()
    torch.zeros(1, device="cuda")  # primary context first: libgstrain.so links the static runtime and shares it
    return dc.torch_backend(C.CDLL(os.path.join(ROOT, "divshot_b200", "lib", "libgstrain.so")))


@pytest.mark.parametrize("case", dc.CASES_PLAIN, ids=lambda c: c.__name__)
def test_refinement_step_on_the_gpu(be, case):
    case(be)


@pytest.mark.parametrize("case", dc.CASES_WITH_OPS, ids=lambda c: c.__name__)
def test_refinement_step_on_the_gpu_vs_per_element_ops(be, ops, case):  # noqa: F811
    case(be, ops)


def test_trainer_refines_through_the_plugin_boundary(tmp_path):
    """600 iterations with warmupLength 100 / refineEvery 100: the MCMC strategy (CLI default) grows the model by 5 % per
    refinement up to capMax; the saved model holds the grown count and finite rows."""
    import subprocess

    import numpy as np
    from divshot_b200 import build
    from test_plugin import LIB, _read_ply
    libs = build.build_all()
    for strategy, expect_growth, extra in (("1", True, []), ("0", None, []), ("1", True, ["visibleAdam=1"]),
                                           ("0", None, ["revisedOpacity=1", "visibleAdam=1"])):
        out = str(tmp_path / f"refined_{strategy}_{len(extra)}.ply")
        r = subprocess.run([libs["gstrain_driver"], "synthetic:N=20000,W=320,H=240,views=4,deg=1", "600", out,
                            "warmup=100", "refineEvery=100", "capMax=24000", "strategy=" + strategy, *extra],
                           capture_output=True, text=True, env={**os.environ, "LD_LIBRARY_PATH": LIB}, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        n, props, rows = _read_ply(out)
        assert np.isfinite(rows).all() and 0 < n <= 24000
        if expect_growth:
            assert n == 24000, f"MCMC: 20000 * 1.05^k capped at capMax, got {n}"


def test_unmodified_reference_cli_refines(tmp_path):
    """The reference's own CLI with its own flags (main.cpp:54-57): warmupLength 100, refineEvery 100, 700 iterations,
    MCMC (its default strategy), capMax 3 000 000 (hard-coded at gs_train.cpp:89): five refinements of +5 %."""
    import subprocess

    import numpy as np
    from divshot_b200 import build
    from test_plugin import LIB, _read_ply
    cli = build.build_all(torch_binding=False).get("reference_cli")
    if not cli:
        pytest.skip("no prebuilt reference CLI")
    out = str(tmp_path / "cli_refined.ply")
    r = subprocess.run([cli, "--inputPath", "synthetic:N=20000,W=320,H=240,views=4,deg=1", "--outputPath", out,
                        "--maxIteration", "700", "--warmupLength", "100", "--refineEvery", "100"], capture_output=True, text=True,
                       env={**os.environ, "LD_LIBRARY_PATH": LIB}, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    n, props, rows = _read_ply(out)
    expect = 20000
    for _ in range(5):  # refinements after iterations 200 .. 600
        expect = int(1.05 * expect)
    assert n == expect and len(props) == 59 and np.isfinite(rows).all()


@pytest.mark.parametrize("W,H,w", [(97, 61, 0.2), (64, 48, 1.0)])
def test_masked_photometric_loss_matches_torch(W, H, w):
    """useMask: loss(M x + (1 - M) y, y) and its gradient w.r.t. x against torch autograd (the unmasked loss kernels
    are GPU-verified by tests/test_plugin.py::test_photometric_loss_matches_torch)."""
    import torch
    import torch.nn.functional as F
    from divshot_b200 import build
    lib = C.CDLL(build.build_gstrain())
    lib.gstrain_masked_photometric_loss.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_float, C.c_void_p]
    lib.gstrain_masked_photometric_loss.restype = None
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu"); g.manual_seed(5)
    x = torch.rand(3, H, W, generator=g).to(dev).requires_grad_(True)
    y = (x.detach() * 0.7 + 0.3 * torch.rand(3, H, W, generator=g).to(dev)).contiguous()
    m = (torch.rand(H, W, generator=g) > 0.4).float()
    m[: H // 3] = torch.rand(H // 3, W, generator=g)  # soft values too
    m = m.to(dev).contiguous()
    k = torch.arange(11, dtype=torch.float64) - 5
    gk = torch.exp(-k ** 2 / (2 * 1.5 ** 2)); gk = (gk / gk.sum()).float().to(dev)
    win = (gk[:, None] * gk[None, :]).expand(3, 1, 11, 11).contiguous()
    conv = lambda t: F.conv2d(t[None], win, padding=5, groups=3)[0]  # noqa: E731
    xb = m[None] * x + (1 - m[None]) * y
    mx, my = conv(xb), conv(y)
    sxx, syy, sxy = conv(xb * xb) - mx * mx, conv(y * y) - my * my, conv(xb * y) - mx * my
    ssim = ((2 * mx * my + 0.01 ** 2) * (2 * sxy + 0.03 ** 2)) / ((mx * mx + my * my + 0.01 ** 2) * (sxx + syy + 0.03 ** 2))
    loss_ref = (1 - w) * (xb - y).abs().mean() + w * (1 - ssim.mean())
    loss_ref.backward()
    xr = x.detach().clone().contiguous()
    dl = torch.empty(3, H, W, device=dev); loss = torch.zeros(1, device=dev); scratch = torch.empty(9 * H * W, device=dev)
    lib.gstrain_masked_photometric_loss(xr.data_ptr(), y.data_ptr(), m.data_ptr(), dl.data_ptr(), loss.data_ptr(), scratch.data_ptr(),
                                        W, H, C.c_float(w), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert float((dl - x.grad).abs().max()) <= 1e-4 * float(x.grad.abs().max()) + 1e-9
    assert float(dl[:, m == 0].abs().max()) == 0.0
