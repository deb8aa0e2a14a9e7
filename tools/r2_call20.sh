#!/usr/bin/env bash
# round 2, GPU call 20-21 (1 GPU): F3 viewer pack after the rework (one resident wave, hoisted reciprocal, float quantiser):
# byte parity against the reference-derived digests + host ops, then the three register budgets
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_viewer_pack.py -m gpu -q -p no:cacheprovider > gpurun_out/c20_tests.log 2>&1
echo "tests exit $?"; tail -6 gpurun_out/c20_tests.log
for c in 8 7 6; do
  DVS_VP_CTAS=$c timeout 300 python tools/bench_viewer_pack.py --steps 50 > gpurun_out/c20_vp_ctas$c.json 2> gpurun_out/c20_vp_ctas$c.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c20_vp_ctas$c.json"))
print("ctas/SM $c:", round(d["ms_per_step"], 4), "ms", round(d["roofline"]["frac"], 3), "of HBM; e2e", round(d["e2e"]["ms_per_step"], 3))
PY
done
