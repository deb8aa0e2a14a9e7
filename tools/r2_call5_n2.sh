#!/usr/bin/env bash
# round 2, GPU call 5 (2 GPUs): the exchanges against the oracle's summed per-view gradients, then the bench at N=2
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/r2_call5_n2.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
tr tools/check_dp_vs_oracle.py --workload c3 --out gpurun_out/c5_dp_vs_oracle_n${N}.json > gpurun_out/c5_dp_check.log 2>&1
echo "dp check exit $?"; grep -E "^rank|Error|error" gpurun_out/c5_dp_check.log | tail -24
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto > gpurun_out/c5_bench_n${N}_auto.json 2> gpurun_out/c5_bench_n${N}_auto.err
echo "bench auto exit $?"; tail -c 500 gpurun_out/c5_bench_n${N}_auto.err
for rc in 16 48 96; do
  tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce fused --fused-reduce-ctas $rc > gpurun_out/c5_bench_n${N}_fused_rc${rc}.json 2> gpurun_out/c5_bench_n${N}_fused_rc${rc}.err
  echo "bench fused rc=$rc exit $?"
done
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/c5_bench_n*.json")):
    try:
        d = json.load(open(p))
        print(p, round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["allreduce"]["backend"], round(d["allreduce"]["ms"], 4), d["allreduce"]["note"][:400])
    except Exception as e:
        print(p, "unreadable", e)
PY
