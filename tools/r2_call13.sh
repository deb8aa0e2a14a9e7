#!/usr/bin/env bash
# round 2, GPU call 13 (1 GPU): first GPU run of the 2DGS path + the other new tests (background, sky, resume)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_2dgs.py -m gpu -q -p no:cacheprovider -x > gpurun_out/c13_2dgs.log 2>&1
echo "2dgs tests exit $?"; tail -25 gpurun_out/c13_2dgs.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_plugin.py -m gpu -q -p no:cacheprovider > gpurun_out/c13_other.log 2>&1
echo "other tests exit $?"; tail -12 gpurun_out/c13_other.log
