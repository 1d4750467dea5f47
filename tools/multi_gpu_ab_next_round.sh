#!/usr/bin/env bash
# Multi-GPU A/B of the exchange variants written after round 1's last GPU session (N = $1, default 2):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/multi_gpu_ab_next_round.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/r02_q3_ab_n$N.log
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@" 2>&1 | tail -2; }
echo "== parity programs over NCCL (pass-through edges now skip the exchange)" | tee $OUT
run tests/dist_nccl_check.py | tee -a $OUT
echo "== ... plus the programs and opt-in plans written after round 1's last GPU session, default and grouped exchange" | tee -a $OUT
run tests/dist_nccl_check.py --more | tee -a $OUT
QK_EXCHANGE=grouped run tests/dist_nccl_check.py --more | tee -a $OUT
echo "== Q3 strong + weak: baseline" | tee -a $OUT
run bench.py --gpus $N --only-q3 | tee -a $OUT
echo "== Q3: replicated build sides (broadcast_cost_based)" | tee -a $OUT
run bench.py --gpus $N --only-q3 --replicate-builds | tee -a $OUT
echo "== Q3: grouped exchange (one batched send/recv call per edge)" | tee -a $OUT
QK_EXCHANGE=grouped run bench.py --gpus $N --only-q3 | tee -a $OUT
echo "== Q3: both" | tee -a $OUT
QK_EXCHANGE=grouped run bench.py --gpus $N --only-q3 --replicate-builds | tee -a $OUT
