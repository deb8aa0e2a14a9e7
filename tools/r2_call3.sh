#!/usr/bin/env bash
# round 2, GPU call 3: baseline of the restored tree — GPU tests, bench line (tight lists), launch list + full ncu capture of one step
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c3_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -6 gpurun_out/c3_gpu_tests.log
timeout 600 python bench.py --no-rows > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench exit $?"; head -c 1500 gpurun_out/c3_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/c3_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-rows > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 80 -c 10 -f -o gpurun_out/c3_step_full \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-rows > /dev/null 2>&1
ls -la gpurun_out | tail -8
