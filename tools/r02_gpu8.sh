#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== exchange vs numpy on 8 ranks" | tee $OUT/r02_g8.log
timeout 300 $TR --master-port 29611 tests/dist_xchg_check.py > $OUT/r02_xchg_n8.log 2>&1; echo "xchg rc=$?" | tee -a $OUT/r02_g8.log
grep -h "OK\|Error" $OUT/r02_xchg_n8.log | tail -3 | tee -a $OUT/r02_g8.log
echo "== DataStream programs on 8 ranks" | tee -a $OUT/r02_g8.log
timeout 600 $TR --master-port 29613 tests/dist_nccl_check.py --more > $OUT/r02_nccl_check_n8.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_g8.log
grep -h "OK\|Error" $OUT/r02_nccl_check_n8.log | tail -3 | tee -a $OUT/r02_g8.log
echo "== full default bench N=8 (the driver's command)" | tee -a $OUT/r02_g8.log
timeout 1500 $TR --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/r02_bench_n8.json 2> $OUT/r02_bench_n8.err; echo "rc=$?" | tee -a $OUT/r02_g8.log
tail -c 2500 $OUT/r02_bench_n8.json | tee -a $OUT/r02_g8.log; grep -v "^\*\|OMP_NUM\|^$" $OUT/r02_bench_n8.err | tail -8 | tee -a $OUT/r02_g8.log
echo "== Q3 only, N=8, shuffle-everything A/B" | tee -a $OUT/r02_g8.log
timeout 300 $TR --master-port 29623 bench.py --gpus 8 --only-q3 --no-cpu --no-replicate-builds 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g8.log
echo done | tee -a $OUT/r02_g8.log
