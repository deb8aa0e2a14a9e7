#!/usr/bin/env bash
# round 2, GPU call 6 (8 GPUs): host-copy probe, exchanges vs the oracle, bench at N=8 (c3/c4 and c5)
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_call6_n8.sh 8'
set -u
N=${1:-8}
mkdir -p gpurun_out
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
nvidia-smi topo -m > gpurun_out/c6_topo.txt 2>&1
tr tools/pcie_probe.py > gpurun_out/c6_pcie_nobind.json 2> gpurun_out/c6_pcie_nobind.err
tr tools/pcie_probe.py --bind > gpurun_out/c6_pcie_bind.json 2> gpurun_out/c6_pcie_bind.err
tail -c 1500 gpurun_out/c6_pcie_nobind.json; echo; tail -c 1500 gpurun_out/c6_pcie_bind.json; echo
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto > gpurun_out/c6_bench_n${N}_auto.json 2> gpurun_out/c6_bench_n${N}_auto.err
echo "bench auto exit $?"; tail -c 300 gpurun_out/c6_bench_n${N}_auto.err
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce fused --fused-reduce-ctas 48 > gpurun_out/c6_bench_n${N}_fused_rc48.json 2> gpurun_out/c6_bench_n${N}_fused_rc48.err
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto --workload c5 > gpurun_out/c6_bench_n${N}_c5.json 2> gpurun_out/c6_bench_n${N}_c5.err
echo "bench c5 exit $?"; tail -c 300 gpurun_out/c6_bench_n${N}_c5.err
tr tools/check_dp_vs_oracle.py --workload c3 --out gpurun_out/c6_dp_vs_oracle_n${N}.json > gpurun_out/c6_dp_check.log 2>&1
echo "dp check exit $?"; grep -E "^rank 0|Error|error" gpurun_out/c6_dp_check.log | tail -8
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/c6_bench_n*.json")):
    try:
        d = json.load(open(p))
        print(p, round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["allreduce"]["backend"], round(d["allreduce"]["ms"], 4), d["config"].get("host_binding"), d["clocks"], d["allreduce"]["note"][:500])
    except Exception as e:
        print(p, "unreadable", e)
PY
