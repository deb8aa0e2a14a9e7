#!/usr/bin/env bash
# round 2, GPU call 15 (1 GPU): 2DGS with sub-tile masks (parity + conservativeness), normal-consistency test, trainer-step bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_2dgs.py -m gpu -q -p no:cacheprovider > gpurun_out/c15_2dgs.log 2>&1
echo "2dgs tests exit $?"; tail -15 gpurun_out/c15_2dgs.log
timeout 600 python -m pytest tests/test_plugin.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "normal_consistency or cull_masks or small_scenes" > gpurun_out/c15_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/c15_tests.log
timeout 400 python tools/bench_trainstep.py > gpurun_out/c15_trainstep.json 2> gpurun_out/c15_trainstep.err
echo "trainstep exit $?"; cat gpurun_out/c15_trainstep.json | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/c15_trainstep.json')); print(d['value'], d['model_2dgs'])"
