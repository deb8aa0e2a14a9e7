"""The editor-facing surface of GaussianTrainerScene (SURVEY.md §8-B "methods used by the editor"): every method the
reference editor calls (application/editor/source/{editor,inspector_panel,scene_view_panel,img2d_dataset_panel}.cpp)
exists with the argument / result types those call sites need.  CPU tier: tools/editor_api_probe.cpp repeats the call
patterns, is compiled with the reference's own glm and linked against libgstrain.so; the class is movable (entt
component) and exported.  STAGED GPU tier (written without a GPU): the probe runs and its outputs are checked."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EDITOR_METHODS = ["loadTrainData", "trainSetup", "trainStep", "saveGaussianModel", "exportMesh", "getNumGaussians",
                  "getGaussianPositionCpu", "getGaussianSH0Cpu", "getGaussianSHNCpu", "getGaussianOpcaitiesCpu",
                  "getGaussianScalingsCpu", "getGaussianRotationsCpu", "getNumCameras", "getCameraProjectionFlat",
                  "getCameraRotationWXYZ", "getCameraPosXYZ", "resetGaussian", "setDensifyStrategy",
                  "getProgressOnCurrentPhase", "getCurrentTrainingPhaseName", "getTrainingElpasedTime",
                  "getEstimateTrainingTime", "updateTensorFromHost", "getPoints3D", "getSplatImageView", "saveCameraDatas",
                  "exportSparsePointCloud", "updateFocusRegion", "getFocusRegionMinMax", "getFocusRegionTransformFlat",
                  "requestViewerPack", "acquireViewerPack"]


def test_every_editor_method_is_exported_and_the_probe_links():
    from divshot_b200 import build
    libs = build.build_all(torch_binding=False)
    syms = subprocess.check_output(["nm", "-DC", "--defined-only", libs["libgstrain"]], text=True)
    for m in EDITOR_METHODS:
        assert f"GaussianTrainerScene::{m}(" in syms or f"GaussianTrainerScene::{m}[abi:cxx11](" in syms, \
            f"libgstrain.so does not export GaussianTrainerScene::{m}"
    hdr = open(os.path.join(ROOT, "include", "gaussian_trainer_scene.hpp")).read()
    for inline in ("getCameraProjection", "getCameraRotation", "getCameraPos", "getFocusRegion", "getFocusRegionTransform",
                   "updateTensorFromGaussianData", "startTrain", "pauseTrain", "isTrain", "isTerminate", "isPruningSplat",
                   "setModelPath", "setTrainingStatus", "getCurrentTrainingStatus", "getCurrentIterations", "maxIteriaons",
                   "getCurrentLoss", "getTrainConfig", "ShowTrainView", "pruenIteraions", "focus_region_position",
                   "focus_region_rotation", "focus_region_scale"):
        assert inline in hdr, inline
    if os.path.isdir("/root/reference/external/glm"):
        assert libs.get("editor_api_probe") and os.path.exists(libs["editor_api_probe"])  # compiled against the reference's glm
        import torch
        if not torch.cuda.is_available():  # no CPU fallback here either
            r = subprocess.run([libs["editor_api_probe"]], capture_output=True, text=True)
            assert r.returncode != 0 and "CUDA" in r.stderr


@pytest.mark.gpu
def test_editor_call_patterns_run(tmp_path):
    from divshot_b200 import build
    exe = build.build_editor_api_probe()
    if not exe:
        pytest.skip("editor_api_probe was not prebuilt (needs the reference's glm at build time)")
    out = str(tmp_path / "probe")
    r = subprocess.run([exe, "synthetic:N=5000,W=160,H=120,views=3,deg=1", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert lines[-1] == "ok" and sum(l.startswith("camera ") for l in lines) == 3
    assert any(l.startswith("edited model 2500 gaussians roundtrip ok") for l in lines)
    assert any(l.startswith("after reset: 5000 gaussians iteration 0") for l in lines)
    assert any(l.startswith("image synthetic_view_1 160x120 alpha 255") for l in lines)
    cams = json.load(open(out + "_cameras.json"))
    assert len(cams) == 3 and cams[1]["width"] == 160 and cams[1]["height"] == 120
    for c in cams:
        R = np.array(c["rotation"])
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-5) and abs(np.linalg.det(R) - 1) < 1e-5
        assert abs(np.hypot(*c["position"][:2]) - 0.5) < 1e-5  # the synthetic ring of radius 0.5
    raw = open(out + "_points.ply", "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert b"element vertex 5000" in head and len(body) == 5000 * 15
    from test_plugin import _read_ply
    n, props, rows = _read_ply(out + ".ply")
    assert n == 5000 and np.isfinite(rows).all()


@pytest.mark.gpu
def test_reference_splatx_cli_end_to_end(tmp_path):
    """The unmodified application/splatx-cli trains through the plugin like diverseshot-cli does (tests/test_plugin.py)."""
    from divshot_b200 import build
    from test_plugin import LIB, _read_ply
    cli = build.build_all(torch_binding=False).get("reference_splatx_cli")
    if not cli:
        pytest.skip("no prebuilt splatx-cli")
    out = str(tmp_path / "splatx.ply")
    r = subprocess.run([cli, "--inputPath", "synthetic:N=20000,W=320,H=240,views=4,deg=1", "--outputPath", out,
                        "--maxIteration", "120"], capture_output=True, text=True,
                       env={**os.environ, "LD_LIBRARY_PATH": LIB}, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    n, props, rows = _read_ply(out)
    assert n == 20000 and len(props) == 59 and np.isfinite(rows).all()
