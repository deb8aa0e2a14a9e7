#!/usr/bin/env bash
# Round 2, first multi-GPU call: the factored gradient exchange against the plain one (same gradients, fewer bytes).
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_multi_gpu_call.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus "$N" --steps 20 --warmup 5 "$@"; }
run --allreduce auto > gpurun_out/bench_n${N}_plain.json 2> gpurun_out/bench_n${N}_plain.err
run --allreduce factored > gpurun_out/bench_n${N}_factored.json 2> gpurun_out/bench_n${N}_factored.err
tail -c 400 gpurun_out/bench_n${N}_plain.json; echo; tail -c 400 gpurun_out/bench_n${N}_factored.json; echo
tail -5 gpurun_out/bench_n${N}_factored.err
