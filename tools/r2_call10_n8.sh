#!/usr/bin/env bash
# round 2, GPU call 10 (8 GPUs): final fused exchange — bench at N=8 (c3/c4 and c5), exchanges vs the oracle
set -u
N=${1:-8}
mkdir -p gpurun_out
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto > gpurun_out/c12_bench_n${N}_auto.json 2> gpurun_out/c12_bench_n${N}_auto.err
echo "bench auto exit $?"; tail -c 300 gpurun_out/c12_bench_n${N}_auto.err
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce fused --fused-reduce-ctas 12 > gpurun_out/c12_bench_n${N}_fused_rc12.json 2> gpurun_out/c12_bench_n${N}_fused_rc12.err
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto --workload c5 > gpurun_out/c12_bench_n${N}_c5.json 2> gpurun_out/c12_bench_n${N}_c5.err
echo "bench c5 exit $?"
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/c12_bench_n*.json")):
    try:
        d = json.load(open(p))
        print(p, round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["allreduce"]["backend"], round(d["allreduce"]["ms"], 4), d["clocks"], d["allreduce"]["note"][:500])
    except Exception as e:
        print(p, "unreadable", e)
PY
