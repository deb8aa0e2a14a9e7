#!/usr/bin/env bash
# round 2, GPU call 1: parity of the rewritten compositing kernels, then A/B against the round-1 build
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c1_parity.log 2>&1
echo "parity exit $?"; tail -15 gpurun_out/c1_parity.log
timeout 900 python tools/ab_bench.py --variants r1 default default+tight nolpt+tight default+tight+absgrad --steps 30 --out gpurun_out/c1_ab.json 2>&1 | tail -20
