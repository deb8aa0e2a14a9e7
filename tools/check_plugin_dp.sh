#!/usr/bin/env bash
# Data parallelism BEHIND the plugin boundary (SURVEY.md §8 e / b): the same unmodified caller loop (tools/gstrain_driver.cpp,
# a stand-in for application/diverseshot-cli/source/gs_train.cpp:105-167) started once per GPU trains one model on W GPUs.
# Compares the model saved by a W-GPU run (one view per rank and step, gradients summed with NCCL inside train_step) with the
# model of a 1-GPU run that accumulates the same W views per step (DVS_BATCH_VIEWS=W), and reports iterations/s of both.
#   /usr/local/graft/bin/gpurun --gpus 2 -- 'bash tools/check_plugin_dp.sh 2'
set -u
W=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export LD_LIBRARY_PATH=$PWD/divshot_b200/lib:${LD_LIBRARY_PATH:-}
DATA="synthetic:N=200000,W=800,H=600,views=8,deg=2"
ITERS=${2:-12}     # parity: a handful of steps (Adam turns rounding noise of near-zero gradients into +-lr moves; over hundreds
                   # of steps those decorrelate element-wise although the loss trajectories stay identical to six digits)
RATE_ITERS=${3:-1000}
DRV=build/gstrain_driver
TMP=$(mktemp -d)
# 1 GPU, W views per step
DVS_BATCH_VIEWS=$W CUDA_VISIBLE_DEVICES=0 $DRV "$DATA" $ITERS $TMP/dp_single.ply lossCheck=0 verbose=0 > $OUT/dp_single.log 2>&1
echo "single exit $?"; tail -2 $OUT/dp_single.log
# W GPUs, one process each (same binary, rank from the environment)
PORT=$((29600 + RANDOM % 300))
for r in $(seq 0 $((W - 1))); do
  DVS_RANK=$r DVS_WORLD_SIZE=$W DVS_LOCAL_RANK=$r MASTER_PORT=$PORT $DRV "$DATA" $ITERS $TMP/dp_multi.ply lossCheck=0 verbose=0 > $OUT/dp_multi_rank$r.log 2>&1 &
done
wait
echo "multi done"; tail -2 $OUT/dp_multi_rank0.log
# iterations/s over a longer run: 1 GPU one view per step, 1 GPU W views per step, W GPUs one view each per step
CUDA_VISIBLE_DEVICES=0 $DRV "$DATA" $RATE_ITERS $TMP/r_plain.ply lossCheck=0 verbose=0 > $OUT/dp_plain.log 2>&1
DVS_BATCH_VIEWS=$W CUDA_VISIBLE_DEVICES=0 $DRV "$DATA" $RATE_ITERS $TMP/r_single.ply lossCheck=0 verbose=0 > $OUT/dp_rate_single.log 2>&1
PORT=$((29600 + RANDOM % 300))
for r in $(seq 0 $((W - 1))); do
  DVS_RANK=$r DVS_WORLD_SIZE=$W DVS_LOCAL_RANK=$r MASTER_PORT=$PORT $DRV "$DATA" $RATE_ITERS $TMP/r_multi.ply lossCheck=0 verbose=0 > $OUT/dp_rate_multi_rank$r.log 2>&1 &
done
wait
tail -1 $OUT/dp_plain.log; tail -1 $OUT/dp_rate_single.log; tail -1 $OUT/dp_rate_multi_rank0.log
python - <<PY
import json, numpy as np
def read(path):
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    n = int([l for l in head.decode().splitlines() if l.startswith("element vertex")][0].split()[-1])
    return np.frombuffer(body, np.float32).reshape(n, -1).astype(np.float64)
a, b = read("$TMP/dp_single.ply"), read("$TMP/dp_multi.ply")
cols = {"means": slice(0, 3), "sh0": slice(3, 6), "shN": slice(6, 51), "opacity": slice(51, 52), "scales": slice(52, 55), "quats": slice(55, 59)}
res = {k: float(np.linalg.norm(a[:, s] - b[:, s]) / (np.linalg.norm(a[:, s]) + 1e-30)) for k, s in cols.items()}
grab = lambda p: [l for l in open(p).read().splitlines() if l.startswith("steps")][-1]
out = {"world": $W, "iters": $ITERS, "data": "$DATA", "norm_rel_diff_multi_vs_single_accumulating": res, "worst": max(res.values()),
       "pass_1e-4": max(res.values()) <= 1e-4, "parity_single_gpu_batch_line": grab("$OUT/dp_single.log"),
       "parity_multi_gpu_rank0_line": grab("$OUT/dp_multi_rank0.log"), "rate_iters": $RATE_ITERS,
       "rate_single_gpu_one_view_per_step": grab("$OUT/dp_plain.log"), "rate_single_gpu_W_views_per_step": grab("$OUT/dp_rate_single.log"),
       "rate_W_gpus_one_view_each_per_step": grab("$OUT/dp_rate_multi_rank0.log")}
json.dump(out, open("$OUT/plugin_dp_n$W.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
rm -rf "$TMP"
