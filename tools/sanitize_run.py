#!/usr/bin/env python
"""Small forward+backward through the C-ABI for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from divshot_b200 import _cabi
from divshot_b200.rasterizer import GradBuffers, Rasterizer, scene_to_device
from divshot_b200.scenes import make_scene

for (N, W, H, deg, shift) in [(3001, 96, 64, 3, 0.8), (517, 37, 21, 1, 1.5), (4000, 64, 64, 0, 2.5)]:
    sc = make_scene(N=N, width=W, height=H, sh_degree=deg, seed=3, normalise_quats=False)
    sc.log_scales += shift
    rast = Rasterizer(0)
    params = scene_to_device(sc, rast.device)
    cam = _cabi.make_camera(sc.cameras[0], deg)
    img, radii = rast.forward(cam, params)
    g = GradBuffers.allocate(sc.N, sc.shN.shape[1], rast.device)
    m2 = torch.zeros(N, 2, device=rast.device); ma = torch.zeros(N, 2, device=rast.device)
    rast.backward(torch.from_numpy(sc.dL_dpix[0]).to(rast.device), g, mean2D=m2, mean2D_abs=ma)
    torch.cuda.synchronize()
    print("ok", N, W, H, deg, rast.stats()["num_dups"], float(img.sum()), float(g.flat.abs().sum()))
    rast.close()
