#!/usr/bin/env bash
# round 2, GPU call 17 (8 GPUs): fused exchange, CTA shape 256 x 2 per SM (default) against 512 x 1; then c5 with the winner's default
set -u
N=${1:-8}
mkdir -p gpurun_out
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto > gpurun_out/c17_bench_n${N}_t256.json 2> gpurun_out/c17_bench_n${N}_t256.err
echo "bench 256x2 exit $?"; tail -c 300 gpurun_out/c17_bench_n${N}_t256.err
DVS_FX_THREADS=512 tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce fused > gpurun_out/c17_bench_n${N}_t512.json 2> gpurun_out/c17_bench_n${N}_t512.err
echo "bench 512x1 exit $?"
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto --workload c5 > gpurun_out/c17_bench_n${N}_c5.json 2> gpurun_out/c17_bench_n${N}_c5.err
echo "bench c5 exit $?"
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/c17_bench_n*.json")):
    try:
        d = json.load(open(p))
        print(p, round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["allreduce"]["backend"], round(d["allreduce"]["ms"], 4), d["clocks"], d["allreduce"]["note"][:400])
    except Exception as e:
        print(p, "unreadable", e)
PY
