#!/usr/bin/env bash
# round 2, GPU call 23 (1 GPU): F3 viewer pack after the second rework: byte parity, 8 vs 6 CTAs per SM, one ncu capture
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_viewer_pack.py -m gpu -q -p no:cacheprovider > gpurun_out/c23_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/c23_tests.log
for c in 8 6; do
  DVS_VP_CTAS=$c timeout 300 python tools/bench_viewer_pack.py --steps 50 > gpurun_out/c23_vp_ctas$c.json 2> gpurun_out/c23_vp_ctas$c.err
  python - <<PY
import json
d = json.load(open("gpurun_out/c23_vp_ctas$c.json"))
print("ctas/SM $c:", round(d["ms_per_step"], 4), "ms", round(d["roofline"]["frac"], 3), "of HBM; e2e", round(d["e2e"]["ms_per_step"], 3))
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viewer_pack_kernel -s 3 -c 1 -f -o gpurun_out/c23_viewer_pack python tools/bench_viewer_pack.py --steps 5 > gpurun_out/c23_vp.log 2>&1
echo "ncu viewer_pack exit $?"
