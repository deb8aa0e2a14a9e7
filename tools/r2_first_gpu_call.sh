#!/usr/bin/env bash
# First GPU call of the next round: run everything that was staged in round 1 without a GPU, then measure it.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_call.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
# 1. the staged tier: refinement step (F1) and viewer hand-off (F3) on the device
timeout 1200 python -m pytest tests -m gpu_staged -q -rf -p no:cacheprovider > gpurun_out/staged_tier.log 2>&1
echo "staged tier exit $?" | tee -a gpurun_out/staged_tier.log
tail -5 gpurun_out/staged_tier.log
# 2. F3 measurement: bench line, launch list, one full capture of the kernel
timeout 600 python tools/bench_viewer_pack.py > gpurun_out/bench_viewer_pack.json 2> gpurun_out/bench_viewer_pack.err
cat gpurun_out/bench_viewer_pack.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/viewer_pack_launches.csv \
    python tools/bench_viewer_pack.py --steps 3 --warmup 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:viewer_pack_kernel -c 1 -f -o gpurun_out/viewer_pack_full \
    python tools/bench_viewer_pack.py --steps 1 --warmup 3 > /dev/null 2>&1
# 2b. F1 measurement
timeout 600 python tools/bench_densify.py > gpurun_out/bench_densify.json 2> gpurun_out/bench_densify.err
cat gpurun_out/bench_densify.json
# 3. the headline bench, unchanged code path (regression check against profiles/r1_final_bench.json: 1.049 ms/step)
timeout 900 python bench.py > gpurun_out/bench_r2_start.json 2> gpurun_out/bench_r2_start.err
tail -c 600 gpurun_out/bench_r2_start.json
