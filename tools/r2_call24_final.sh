#!/usr/bin/env bash
# round 2, GPU call 24 (1 GPU), final validation: the whole gpu tier on the final build, the A/B switches' other settings,
# the bench line with its rows, the launch list and one ncu --set full capture of a step, the c5 line
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c24_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -4 gpurun_out/c24_gpu_tests.log
DVS_VP_PREFETCH=0 timeout 300 python -m pytest tests/test_viewer_pack.py -m gpu -q -p no:cacheprovider 2>&1 | tail -1
DVS_VP_CTAS=8 timeout 300 python -m pytest tests/test_viewer_pack.py -m gpu -q -p no:cacheprovider 2>&1 | tail -1
DVS_SURFEL_PB_DIRECT=1 timeout 300 python -m pytest tests/test_gpu_2dgs.py -m gpu -q -p no:cacheprovider 2>&1 | tail -1
for v in "1 6" "1 5" "1 8" "0 6"; do
  set -- $v
  DVS_VP_PREFETCH=$1 DVS_VP_CTAS=$2 timeout 300 python tools/bench_viewer_pack.py --steps 50 > gpurun_out/c24_vp_p$1_c$2.json 2> /dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/c24_vp_p$1_c$2.json"))
print("viewer pack prefetch $1 ctas/SM $2:", round(d["ms_per_step"], 4), "ms", round(d["roofline"]["frac"], 3), "of HBM")
PY
done
DVS_SURFEL_PB_DIRECT=1 timeout 300 python tools/ab_bench.py --variants default+2dgs --steps 20 --out gpurun_out/c24_ab_2dgs_direct.json 2>&1 | tail -1 | cut -c1-330
timeout 300 python tools/ab_bench.py --variants default+2dgs --steps 20 --out gpurun_out/c24_ab_2dgs.json 2>&1 | tail -1 | cut -c1-330
timeout 900 python bench.py > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/c24_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["ms_per_step"], {k: v["ms"] for k, v in d["stages"].items()}, d["roofline_step"])
print({k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in d["other_rows"].items()})
PY
timeout 600 python bench.py --workload c5 --no-rows --no-cpu > gpurun_out/c24_bench_c5.json 2> gpurun_out/c24_bench_c5.err
echo "bench c5 exit $?"; python -c "
import json; d = json.load(open('gpurun_out/c24_bench_c5.json')); print('c5', d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/c24_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-rows > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 80 -c 14 -f -o gpurun_out/c24_step_full \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-rows > /dev/null 2>&1
ls -la gpurun_out | grep c24 | awk '{print $5, $9}'
