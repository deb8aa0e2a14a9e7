#!/usr/bin/env bash
# round 2, GPU call 25 (1 GPU), final build with the cluster tile scan: gpu tier, A/B against the single-CTA scan, step time,
# launch list + ncu --set full capture, then the complete bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/c25_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 gpurun_out/c25_gpu_tests.log
DVS_SCAN_SINGLE_CTA=1 timeout 200 python tools/ab_bench.py --variants default+tight --steps 30 --out gpurun_out/c25_ab_scan_single.json 2>&1 | tail -1 | cut -c1-300
timeout 200 python tools/ab_bench.py --variants default+tight --steps 30 --out gpurun_out/c25_ab_scan_cluster.json 2>&1 | tail -1 | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/c25_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-rows > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -s 80 -c 14 -f -o gpurun_out/c25_step_full \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-rows > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/c25_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["ms_per_step"], {k: v["ms"] for k, v in d["stages"].items()}, d["roofline_step"])
print({k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in d["other_rows"].items()})
PY
ls -la gpurun_out | grep c25 | awk '{print $5, $9}'
