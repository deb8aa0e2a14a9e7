"""Runs the STAGED GPU tier (marker `gpu_staged`: tests of code written while this round had no GPU minutes left —
the refinement step, the viewer hand-off, the editor surface, the auxiliary maps) once, near the end of the `-m gpu` run, in a
SUBPROCESS:
  * a crash or a sticky CUDA error in unproven code cannot take the proven tests down with it (separate process);
  * the outcome is visible either way — this test passes when every staged test passes, and reports `xfailed` with the
    failing test names otherwise (nothing staged is claimed as GPU-verified in DESIGN.md until it passes here)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_staged_gpu_tier_first_run():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu_staged", "-q", "-p", "no:cacheprovider",
                        "--no-header", "-rf"], capture_output=True, text=True, cwd=ROOT, timeout=1500)
    tail = "\n".join(r.stdout.strip().splitlines()[-25:])
    m = re.search(r"(\d+) passed", r.stdout)
    failed = re.search(r"(\d+) (failed|error)", r.stdout)
    if r.returncode == 0 and m and not failed:
        print(f"staged GPU tier: {m.group(1)} passed")
        return
    pytest.xfail("staged GPU tier did not pass on its first GPU run:\n" + tail + "\n" + r.stderr[-1500:])
