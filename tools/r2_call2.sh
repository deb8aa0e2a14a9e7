#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c2_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -8 gpurun_out/c2_gpu_tests.log
timeout 600 python tools/ab_bench.py --variants r1 default+tight default --steps 30 --out gpurun_out/c2_ab.json 2>&1 | tail -20
