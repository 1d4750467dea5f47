#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tests/dist_xchg_check.py > $OUT/r02_xchg.log 2>&1; echo "xchg rc=$?" | tee $OUT/r02_g3.log
QK_MAILBOX_MB=1 timeout 300 $TR --master-port 29612 tests/dist_xchg_check.py > $OUT/r02_xchg_small.log 2>&1; echo "xchg small rc=$?" | tee -a $OUT/r02_g3.log
timeout 600 $TR --master-port 29613 tests/dist_nccl_check.py --more > $OUT/r02_nccl_check.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_g3.log
grep -h "OK\|Error" $OUT/r02_xchg.log $OUT/r02_xchg_small.log $OUT/r02_nccl_check.log | tail -12 | tee -a $OUT/r02_g3.log
for cr in 0 134217728 67108864 33554432; do
  echo "== Q3 N=1 chunk_rows=$cr" | tee -a $OUT/r02_g3.log
  timeout 300 python bench.py --only-q3 --no-cpu --chunk-rows $cr 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g3.log
done
for cr in 0 67108864 33554432; do
  echo "== Q3 N=2 chunk_rows=$cr" | tee -a $OUT/r02_g3.log
  timeout 300 $TR --master-port 29620 bench.py --gpus 2 --only-q3 --no-cpu --chunk-rows $cr 2>&1 | tail -1 | cut -c1-420 | tee -a $OUT/r02_g3.log
done
echo "== asof N=2" | tee -a $OUT/r02_g3.log
timeout 300 $TR --master-port 29621 bench.py --gpus 2 --only-asof --no-cpu 2>&1 | tail -1 | tee -a $OUT/r02_g3.log
