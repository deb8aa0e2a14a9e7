#!/usr/bin/env bash
# round 2, GPU call 8 (2 GPUs): pull-based fused exchange vs the oracle, bench N=2, data parallelism behind the plugin, e2e probe
set -u
N=2
mkdir -p gpurun_out
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
tr tools/check_dp_vs_oracle.py --workload c3 --out gpurun_out/c8_dp_vs_oracle_n${N}.json > gpurun_out/c8_dp_check.log 2>&1
echo "dp check exit $?"; grep -E "^rank 0|Error|error" gpurun_out/c8_dp_check.log | tail -12
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto > gpurun_out/c8_bench_n${N}_auto.json 2> gpurun_out/c8_bench_n${N}_auto.err
echo "bench auto exit $?"; tail -c 400 gpurun_out/c8_bench_n${N}_auto.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c8_bench_n2_auto.json"))
print(round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["allreduce"]["backend"], round(d["allreduce"]["ms"], 4), d["allreduce"]["note"][:600])
PY
timeout 900 bash tools/check_plugin_dp.sh 2 2>&1 | tail -30

