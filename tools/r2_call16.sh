#!/usr/bin/env bash
# round 2, GPU call 16 (1 GPU): stage times of the 2DGS variant at c3, dataset-loader / packed-image test
set -u
mkdir -p gpurun_out
timeout 600 python tools/ab_bench.py --variants default+tight default+2dgs --steps 20 --out gpurun_out/c16_ab_2dgs.json 2>&1 | tail -3
timeout 600 python -m pytest tests/test_plugin.py -m gpu -q -p no:cacheprovider -k "dataset_directory" > gpurun_out/c16_tests.log 2>&1
echo "tests exit $?"; tail -8 gpurun_out/c16_tests.log
