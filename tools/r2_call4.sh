#!/usr/bin/env bash
# round 2, GPU call 4 (1 GPU): the whole gpu tier incl. the promoted and the new tests, then the bench line (pipelined e2e, full-c3 CPU arm)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/c4_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -12 gpurun_out/c4_gpu_tests.log
timeout 600 python bench.py > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
echo "bench exit $?"; tail -c 300 gpurun_out/c4_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/c4_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "cpu_baseline", "clocks")})
print(d["roofline"]); print(d.get("other_rows"))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c4_bench_ref.json 2> gpurun_out/c4_bench_ref.err
tail -c 700 gpurun_out/c4_bench_ref.json
