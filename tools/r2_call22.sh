#!/usr/bin/env bash
# round 2, GPU call 22 (1 GPU): 2DGS with ballot-driven entry walks (parity + stage times); ncu --set full of the reworked
# viewer_pack_kernel
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_2dgs.py -m gpu -q -p no:cacheprovider > gpurun_out/c22_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/c22_tests.log
timeout 600 python tools/ab_bench.py --variants default+2dgs --steps 20 --out gpurun_out/c22_ab_2dgs.json 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:viewer_pack_kernel -s 3 -c 1 -f -o gpurun_out/c22_viewer_pack python tools/bench_viewer_pack.py > gpurun_out/c22_vp.log 2>&1
echo "ncu viewer_pack exit $?"
