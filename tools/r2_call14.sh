#!/usr/bin/env bash
# round 2, GPU call 14 (1 GPU): normal-consistency loss test, trainer-step bench (3DGS + 2DGS), full bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_plugin.py -m gpu -q -p no:cacheprovider -k "normal_consistency or sky or fused_adam" > gpurun_out/c14_tests.log 2>&1
echo "tests exit $?"; tail -6 gpurun_out/c14_tests.log
timeout 400 python tools/bench_trainstep.py > gpurun_out/c14_trainstep.json 2> gpurun_out/c14_trainstep.err
echo "trainstep exit $?"; cat gpurun_out/c14_trainstep.json; tail -3 gpurun_out/c14_trainstep.err
timeout 900 python bench.py > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/c14_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["ms_per_step"], {k: v["ms"] for k, v in d["stages"].items()}, d["roofline_step"])
print({k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in d["other_rows"].items()})
PY
