#!/usr/bin/env bash
# Round 2, GPU call 2 (2 GPUs): the peer-memory exchange against numpy, the DataStream programs on 2 ranks, Q3 strong/weak.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 1. exchange vs numpy" | tee $OUT/r02_g2.log
timeout 300 $TR --master-port 29611 tests/dist_xchg_check.py 2>&1 | tail -25 | tee -a $OUT/r02_g2.log
echo "== 1b. exchange, 1 MB mailbox (rounds)" | tee -a $OUT/r02_g2.log
QK_MAILBOX_MB=1 timeout 300 $TR --master-port 29612 tests/dist_xchg_check.py 2>&1 | tail -25 | tee -a $OUT/r02_g2.log
echo "== 2. DataStream programs on 2 ranks" | tee -a $OUT/r02_g2.log
timeout 600 $TR --master-port 29613 tests/dist_nccl_check.py --more 2>&1 | tail -40 | tee -a $OUT/r02_g2.log
echo "== 3. Q3 SF-100 strong + weak, 2 GPUs (peer-memory exchange)" | tee -a $OUT/r02_g2.log
timeout 600 $TR --master-port 29614 bench.py --gpus 2 --only-q3 --no-cpu 2>&1 | tail -3 | tee $OUT/r02_q3_n2.json | cut -c1-2500 | tee -a $OUT/r02_g2.log
echo "== 3b. same over NCCL (QK_P2P=0)" | tee -a $OUT/r02_g2.log
QK_P2P=0 timeout 600 $TR --master-port 29615 bench.py --gpus 2 --only-q3 --no-cpu 2>&1 | tail -3 | tee $OUT/r02_q3_n2_nccl.json | cut -c1-1500 | tee -a $OUT/r02_g2.log
echo "== 4. the one failing test of call 1" | tee -a $OUT/r02_g2.log
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 | tee -a $OUT/r02_g2.log
echo done | tee -a $OUT/r02_g2.log
