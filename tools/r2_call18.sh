#!/usr/bin/env bash
# round 2, GPU call 18 (1 GPU): 2DGS after the compositor rework (reciprocals, warp-uniform forward walk, recursive-halving
# warp sums, exact projected-ellipse cull): parity tests + stage times at c3
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_2dgs.py -m gpu -q -p no:cacheprovider > gpurun_out/c18_tests.log 2>&1
echo "tests exit $?"; tail -8 gpurun_out/c18_tests.log
timeout 600 python tools/ab_bench.py --variants default+2dgs --steps 20 --out gpurun_out/c18_ab_2dgs.json 2>&1 | tail -3
