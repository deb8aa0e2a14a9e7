#!/usr/bin/env bash
# round 2, GPU call 7 (1 GPU): gpu tier (new tests), A/B of the compositing-backward / emission micro-optimisations, bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c7_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -8 gpurun_out/c7_gpu_tests.log
timeout 600 python tools/ab_bench.py --variants r2a+tight default+tight --steps 30 --out gpurun_out/c7_ab.json 2>&1 | tail -4
timeout 600 python bench.py --no-rows > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c7_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["ms_per_step"], {k: v["ms"] for k, v in d["stages"].items()})
PY
timeout 300 python tools/bench_densify.py > gpurun_out/c7_densify.json 2>&1; tail -c 400 gpurun_out/c7_densify.json
