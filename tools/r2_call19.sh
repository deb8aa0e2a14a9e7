#!/usr/bin/env bash
# round 2, GPU call 19 (1 GPU): ncu --set full of viewer_pack_kernel (F3) and of the surfel compositing backward
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viewer_pack_kernel -s 3 -c 1 -f -o gpurun_out/c19_viewer_pack python tools/bench_viewer_pack.py > gpurun_out/c19_vp.log 2>&1
echo "ncu viewer_pack exit $?"; tail -2 gpurun_out/c19_vp.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:surfel_render -s 4 -c 2 -f -o gpurun_out/c19_surfel python tools/ab_bench.py --variants default+2dgs --steps 3 --out gpurun_out/c19_ab.json > gpurun_out/c19_sf.log 2>&1
echo "ncu surfel exit $?"; tail -2 gpurun_out/c19_sf.log
ls -la gpurun_out/*.ncu-rep
