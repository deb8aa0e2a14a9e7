#!/usr/bin/env bash
# round 2, GPU call 11 (2 GPUs): cp.async-staged fused exchange sanity + bench; plugin DP exchange timing / NCCL transport
set -u
N=2
mkdir -p gpurun_out
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
tr tools/check_dp_vs_oracle.py --workload c3 --out gpurun_out/c11_dp_vs_oracle_n${N}.json > gpurun_out/c11_dp_check.log 2>&1
echo "dp check exit $?"; grep -E "^rank 0|Error|error" gpurun_out/c11_dp_check.log | tail -12
tr bench.py --gpus "$N" --steps 20 --warmup 5 --allreduce auto > gpurun_out/c11_bench_n${N}_auto.json 2> gpurun_out/c11_bench_n${N}_auto.err
echo "bench auto exit $?"; tail -c 300 gpurun_out/c11_bench_n${N}_auto.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c11_bench_n2_auto.json"))
print(round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["allreduce"]["backend"], round(d["allreduce"]["ms"], 4), d["allreduce"]["note"][:600])
PY
# plugin DP: where does the time go?  system NCCL vs the NCCL torch bundles, with the exchange timed on the device
export LD_LIBRARY_PATH=$PWD/divshot_b200/lib:${LD_LIBRARY_PATH:-}
DATA="synthetic:N=200000,W=800,H=600,views=8,deg=2"
TORCH_NCCL=$(python -c "import nvidia.nccl, os; print(os.path.join(os.path.dirname(nvidia.nccl.__file__), 'lib'))" 2>/dev/null)
for variant in system torch; do
  PORT=$((29600 + RANDOM % 300))
  for r in 0 1; do
    if [ "$variant" = torch ] && [ -n "$TORCH_NCCL" ]; then EXTRA="$TORCH_NCCL:"; else EXTRA=""; fi
    LD_LIBRARY_PATH="$EXTRA$LD_LIBRARY_PATH" NCCL_DEBUG=INFO DVS_DP_TIMING=1 DVS_RANK=$r DVS_WORLD_SIZE=2 DVS_LOCAL_RANK=$r MASTER_PORT=$PORT \
      build/gstrain_driver "$DATA" 600 /tmp/dp_dbg_$variant.ply lossCheck=0 verbose=0 > gpurun_out/c11_dp_${variant}_rank$r.log 2>&1 &
  done
  wait
  echo "== $variant NCCL"; grep -E "its_per_s|gradient exchange|NCCL version|via P2P|via SHM|via NET|NVLS|Connected all" gpurun_out/c11_dp_${variant}_rank0.log | head -12
done
CUDA_VISIBLE_DEVICES=0 build/gstrain_driver "$DATA" 600 /tmp/dp_dbg_single.ply lossCheck=0 verbose=0 2>&1 | tail -1
CUDA_VISIBLE_DEVICES=0 timeout 120 python tools/e2e_probe.py 2>/dev/null | tail -1
timeout 300 bash tools/check_plugin_dp.sh 2 4 300 2>&1 | grep -E "worst|pass|means|sh0|opacity|scales|quats"
