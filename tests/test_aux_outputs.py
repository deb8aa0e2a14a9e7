"""SURVEY.md §8 row F4 — depth / alpha maps and their gradients (for depth- and normal-consistency losses, mesh
extraction), produced by linearity from the existing compositor (tests/aux_ref.py explains the construction).
CPU tier: the oracle-level restatement against a float64 autograd re-expression that shares no code with it.
GPU tier (STAGED): dvs_rast_forward_aux / dvs_rast_backward_aux against the restatement."""
import numpy as np
import pytest

import autograd_ref as ar
import aux_ref
from divshot_b200.scenes import make_scene
from oracle import oracle as orc
from util import assert_close, assert_close_robust, orc_cam, scene_arrays


def _scene(seed, deg, bg=(0, 0, 0)):
    sc = make_scene(N=1200, width=64, height=48, sh_degree=deg, seed=seed, normalise_quats=False, bg=bg)
    sc.log_scales += 1.2
    return sc


@pytest.mark.parametrize("deg,seed,bg", [(1, 11, (0, 0, 0)), (3, 12, (0.3, 0.1, 0.7))])
def test_oracle_restatement_matches_float64_autograd(deg, seed, bg):
    sc = _scene(seed, deg, bg)
    cam = sc.cameras[0]
    oc = orc_cam(cam, deg)
    arrays = scene_arrays(sc)
    fwd = orc.forward(oc, *arrays, threads=1)
    rng = np.random.default_rng(seed)
    dL_daux = rng.normal(size=(2, cam.height, cam.width)).astype(np.float32)
    dL_daux[0] *= 0.2
    img, g, proj, _, final_T = ar.render_and_grad(cam, arrays, deg, fwd.ranges, fwd.point_list, fwd.radii, sc.dL_dpix[0], dL_daux=dL_daux)
    aux = aux_ref.forward_aux(oc, fwd)
    assert_close(aux[0], proj["aux"][0], 1e-4, "depth map")
    assert_close(aux[1], proj["aux"][1], 1e-4, "alpha map")
    assert np.allclose(aux[1], 1.0 - fwd.final_T.reshape(cam.height, cam.width), atol=2e-6), "alpha = 1 - T"
    assert aux[0].max() > 2.0 and (aux[0] >= 0).all()  # depths of this scene are 2..10
    b = aux_ref.backward_with_aux(oc, fwd, *arrays, sc.dL_dpix[0], dL_daux)
    assert_close_robust(b["means3D"], g["means3D"], 1e-4, "dL_dmeans3D", frac=0.995)
    assert_close_robust(b["scales"], g["scales"], 1e-4, "dL_dscales", frac=0.995)
    assert_close_robust(b["quats"], g["quats"], 1e-4, "dL_dquats", frac=0.995)
    assert_close_robust(b["opac"], g["opac"].reshape(-1), 1e-4, "dL_dopacity", frac=0.995)
    assert_close(b["sh0"], g["sh0"], 2e-4, "dL_dsh0")
    assert_close(b["shN"], g["shN"], 2e-4, "dL_dshN")
    # the depth term really matters in this test: without it the mean gradient is visibly different
    plain = orc.backward(oc, fwd, *arrays, sc.dL_dpix[0], threads=1)
    assert np.abs(plain.dL_dmeans3D - b["means3D"]).max() > 1e-2 * np.abs(b["means3D"]).max()


@pytest.mark.parametrize("deg,seed,normq", [(1, 21, False), (2, 22, True)])
def test_normal_map_restatement_matches_float64_autograd(deg, seed, normq):
    sc = _scene(seed, deg)
    if normq:
        sc.quats /= np.linalg.norm(sc.quats, axis=1, keepdims=True)
    else:
        sc.quats *= np.random.default_rng(seed).uniform(0.3, 3.0, (sc.N, 1)).astype(np.float32)  # the kernel normalises
    cam = sc.cameras[0]
    oc = orc_cam(cam, deg)
    arrays = scene_arrays(sc)
    fwd = orc.forward(oc, *arrays, threads=1)
    rng = np.random.default_rng(seed)
    dL_d5 = rng.normal(size=(5, cam.height, cam.width)).astype(np.float32)
    dL_d5[0] *= 0.2
    img, g, proj, _, _ = ar.render_and_grad(cam, arrays, deg, fwd.ranges, fwd.point_list, fwd.radii, sc.dL_dpix[0], dL_daux=dL_d5)
    nmap = aux_ref.forward_normals(oc, fwd, *arrays[:3])
    assert_close(nmap, proj["aux"][2:5], 1e-4, "normal map")
    n_v, axis, flip = aux_ref.normals(oc, *arrays[:3])
    vis = fwd.radii > 0
    assert np.allclose(np.linalg.norm(n_v[vis], axis=1), 1.0, atol=1e-5) and set(np.unique(axis)) <= {0, 1, 2}
    assert np.allclose(n_v[vis], proj["normal"].detach().numpy()[vis], atol=2e-6)
    b = aux_ref.backward_with_aux(oc, fwd, *arrays, sc.dL_dpix[0], dL_d5[:2], dL_d5[2:5])
    assert_close_robust(b["quats"], g["quats"], 1e-4, "dL_dquats", frac=0.995)
    assert_close_robust(b["means3D"], g["means3D"], 1e-4, "dL_dmeans3D", frac=0.995)
    assert_close_robust(b["scales"], g["scales"], 1e-4, "dL_dscales", frac=0.995)
    assert_close_robust(b["opac"], g["opac"].reshape(-1), 1e-4, "dL_dopacity", frac=0.995)
    # the normal term really reaches the rotations
    plain = aux_ref.backward_with_aux(oc, fwd, *arrays, sc.dL_dpix[0], dL_d5[:2])
    assert np.abs(plain["quats"] - b["quats"]).max() > 1e-2 * np.abs(b["quats"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("deg,seed,bg,N,W,H", [(1, 11, (0, 0, 0), 1200, 64, 48), (3, 12, (0.3, 0.1, 0.7), 4000, 128, 96),
                                                (2, 13, (1, 1, 1), 30000, 320, 200)])
def test_cuda_aux_outputs_match_the_oracle_restatement(deg, seed, bg, N, W, H):
    import torch
    from divshot_b200 import _cabi
    from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
    sc = make_scene(N=N, width=W, height=H, sh_degree=deg, seed=seed, normalise_quats=False, bg=bg)
    sc.log_scales += 0.8
    cam = sc.cameras[0]
    oc = orc_cam(cam, deg)
    arrays = scene_arrays(sc)
    fwd = orc.forward(oc, *arrays)
    rng = np.random.default_rng(seed)
    dL_daux = rng.normal(size=(2, H, W)).astype(np.float32)
    dL_daux[0] *= 0.2
    dL_dn = rng.normal(size=(3, H, W)).astype(np.float32)
    r = Rasterizer(0)
    try:
        params = scene_to_device(sc, r.device)
        dcam = _cabi.make_camera(cam, deg)
        img, radii = r.forward(dcam, params)
        aux, nmap = r.forward_aux(normals=True)
        aux, nmap = aux.cpu().numpy(), nmap.cpu().numpy()
        ref = aux_ref.forward_aux(oc, fwd)
        assert_close(aux[0], ref[0], 1e-4, "depth map")
        assert_close(aux[1], ref[1], 1e-4, "alpha map")
        assert_close(nmap, aux_ref.forward_normals(oc, fwd, *arrays[:3]), 1e-4, "normal map")
        assert np.array_equal(r.forward_aux().cpu().numpy(), aux)  # depth / alpha alone, repeatable
        g = GradBuffers.allocate(sc.N, sc.shN.shape[1], r.device)
        g.flat.fill_(float("nan"))
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(r.device)  # noqa: E731
        r.backward_aux(dev(sc.dL_dpix[0]), dev(dL_daux), g, dL_dnormal=dev(dL_dn))
        torch.cuda.synchronize()
        b = aux_ref.backward_with_aux(oc, fwd, *arrays, sc.dL_dpix[0], dL_daux, dL_dn)
        assert_close_robust(g.means3D.cpu().numpy(), b["means3D"], 1e-4, "dL_dmeans3D", frac=0.995)
        assert_close_robust(g.scales.cpu().numpy(), b["scales"], 1e-4, "dL_dscales", frac=0.995)
        assert_close_robust(g.quats.cpu().numpy(), b["quats"], 1e-4, "dL_dquats", frac=0.995)
        assert_close_robust(g.opacities.cpu().numpy().reshape(-1), b["opac"], 1e-4, "dL_dopacity", frac=0.995)
        assert_close(g.sh0.cpu().numpy(), b["sh0"], 2e-4, "dL_dsh0")
        assert_close(g.shN.cpu().numpy(), b["shN"], 2e-4, "dL_dshN")
        # the plain backward afterwards is unaffected (the screen-gradient records were left clean)
        g2 = GradBuffers.allocate(sc.N, sc.shN.shape[1], r.device)
        r.forward(dcam, params)
        r.backward(dev(sc.dL_dpix[0]), g2)
        torch.cuda.synchronize()
        plain = orc.backward(oc, fwd, *arrays, sc.dL_dpix[0])
        assert_close_robust(g2.means3D.cpu().numpy(), plain.dL_dmeans3D, 1e-4, "plain dL_dmeans3D", frac=0.995)
        assert_close(g2.sh0.cpu().numpy(), plain.dL_dsh0, 2e-4, "plain dL_dsh0")
        # zero auxiliary gradients == plain backward; each auxiliary loss can be given alone
        g3 = GradBuffers.allocate(sc.N, sc.shN.shape[1], r.device)
        r.backward_aux(dev(sc.dL_dpix[0]), torch.zeros(2, H, W, device=r.device), g3, dL_dnormal=torch.zeros(3, H, W, device=r.device))
        torch.cuda.synchronize()
        assert_close_robust(g3.means3D.cpu().numpy(), g2.means3D.cpu().numpy(), 5e-5, "zero-aux means", frac=0.99)  # atomics: order-dependent rounding only
        assert_close_robust(g3.quats.cpu().numpy(), g2.quats.cpu().numpy(), 5e-5, "zero-aux quats", frac=0.99)
        g4 = GradBuffers.allocate(sc.N, sc.shN.shape[1], r.device)
        r.backward_aux(dev(sc.dL_dpix[0]), None, g4, dL_dnormal=dev(dL_dn))
        torch.cuda.synchronize()
        b4 = aux_ref.backward_with_aux(oc, fwd, *arrays, sc.dL_dpix[0], np.zeros((2, H, W), np.float32), dL_dn)
        assert_close_robust(g4.quats.cpu().numpy(), b4["quats"], 1e-4, "normal-only dL_dquats", frac=0.995)
    finally:
        r.close()


@pytest.mark.gpu
def test_libtorch_rasterize_aux_matches_c_abi():
    """torch.ops.dvs.rasterize_aux (csrc/torch_binding.cpp): image, radii, depth/alpha and normal maps with autograd through
    dvs_rast_backward_aux — against the Python / C-ABI path on the same scene."""
    import torch
    from divshot_b200 import _cabi, build
    from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
    torch.classes.load_library(build.build_all()["libdvs_torch"])
    sc = make_scene(N=5000, width=128, height=96, sh_degree=2, seed=5)
    sc.log_scales += 0.8
    dev = torch.device("cuda", 0)
    params = scene_to_device(sc, dev)
    cam = sc.cameras[0]
    packed = np.zeros(48, np.float32)
    packed[0:16] = cam.view; packed[16:32] = cam.proj; packed[32:35] = cam.campos
    packed[35:37] = (cam.tanfovx, cam.tanfovy); packed[37:39] = (cam.width, cam.height); packed[39:42] = cam.bg
    packed[42:46] = (1.0, 2, 8, 0)
    r = torch.classes.dvs.Rasterizer(0)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    img, radii, aux, normal = torch.ops.dvs.rasterize_aux(r, torch.from_numpy(packed), leaves["means3D"], leaves["scales"], leaves["quats"],
                                                          leaves["opacities"], leaves["sh0"], leaves["shN"])
    g = torch.Generator(device="cpu"); g.manual_seed(1)
    dl = torch.from_numpy(sc.dL_dpix[0]).to(dev)
    da = (0.3 * torch.randn(2, cam.height, cam.width, generator=g)).to(dev)
    dn = torch.randn(3, cam.height, cam.width, generator=g).to(dev)
    ((img * dl).sum() + (aux * da).sum() + (normal * dn).sum()).backward()
    ref = Rasterizer(0)
    try:
        rimg, rradii = ref.forward(_cabi.make_camera(cam, 2), params)
        raux, rnormal = ref.forward_aux(normals=True)
        gb = GradBuffers.allocate(sc.N, 8, dev)
        ref.backward_aux(dl, da, gb, dL_dnormal=dn)
        torch.cuda.synchronize()
        assert torch.equal(img.detach(), rimg) and torch.equal(radii, rradii)
        assert torch.equal(aux.detach(), raux) and torch.equal(normal.detach(), rnormal)
        for k in ("means3D", "scales", "quats", "opacities", "sh0", "shN"):
            assert_close_robust(leaves[k].grad.cpu().numpy(), getattr(gb, k).cpu().numpy(), 1e-4, k)
        # only the colour loss: rasterize_aux degrades to rasterize
        leaves2 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        img2, _, _, _ = torch.ops.dvs.rasterize_aux(r, torch.from_numpy(packed), leaves2["means3D"], leaves2["scales"], leaves2["quats"],
                                                    leaves2["opacities"], leaves2["sh0"], leaves2["shN"])
        (img2 * dl).sum().backward()
        gp = GradBuffers.allocate(sc.N, 8, dev)
        ref.forward(_cabi.make_camera(cam, 2), params)
        ref.backward(dl, gp)
        torch.cuda.synchronize()
        assert_close_robust(leaves2["means3D"].grad.cpu().numpy(), gp.means3D.cpu().numpy(), 1e-4, "colour-only means3D")
    finally:
        ref.close()
