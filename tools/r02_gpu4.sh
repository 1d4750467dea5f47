#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
QK_MAILBOX_MB=1 timeout 300 $TR --master-port 29612 tests/dist_xchg_check.py > $OUT/r02_xchg_small.log 2>&1; echo "xchg small rc=$?" | tee $OUT/r02_g4.log
grep -h "OK\|Error" $OUT/r02_xchg_small.log | tail -4 | tee -a $OUT/r02_g4.log
echo "== Q3 N=2 cProfile" | tee -a $OUT/r02_g4.log
QK_CPROFILE=$OUT/r02_q3_n2_cprofile.txt timeout 300 $TR --master-port 29620 bench.py --gpus 2 --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-300 | tee -a $OUT/r02_g4.log
echo "== Q3 N=1 cProfile" | tee -a $OUT/r02_g4.log
QK_CPROFILE=$OUT/r02_q3_n1_cprofile.txt timeout 300 python bench.py --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-300 | tee -a $OUT/r02_g4.log
for cr in 134217728 67108864; do
  echo "== Q3 N=1 chunk_rows=$cr" | tee -a $OUT/r02_g4.log
  timeout 300 python bench.py --only-q3 --no-cpu --chunk-rows $cr 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g4.log
  echo "== Q3 N=2 chunk_rows=$cr" | tee -a $OUT/r02_g4.log
  timeout 300 $TR --master-port 29621 bench.py --gpus 2 --only-q3 --no-cpu --chunk-rows $cr 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g4.log
done
echo "== asof N=2 (200M quotes/GPU) with QK_PROFILE" | tee -a $OUT/r02_g4.log
QK_PROFILE=1 timeout 300 $TR --master-port 29622 bench.py --gpus 2 --only-asof --no-cpu --asof-quotes 200000000 2>&1 | grep -v "^\*\|OMP\|NCCL" | tail -3 | tee -a $OUT/r02_g4.log
echo "== asof N=1 default scale (1.05B quotes)" | tee -a $OUT/r02_g4.log
timeout 300 python bench.py --only-asof --no-cpu 2>&1 | tail -1 | tee -a $OUT/r02_g4.log
