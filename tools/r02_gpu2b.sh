#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tests/dist_xchg_check.py > $OUT/r02_xchg.log 2>&1; echo "xchg rc=$?" | tee $OUT/r02_g2b.log
QK_MAILBOX_MB=1 timeout 300 $TR --master-port 29612 tests/dist_xchg_check.py > $OUT/r02_xchg_small.log 2>&1; echo "xchg small rc=$?" | tee -a $OUT/r02_g2b.log
timeout 600 $TR --master-port 29613 tests/dist_nccl_check.py --more > $OUT/r02_nccl_check.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_g2b.log
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -x -p no:cacheprovider -k "asof or compact or bloom or partition" 2>&1 | tail -15 | tee -a $OUT/r02_g2b.log
timeout 300 python bench.py --only-asof --no-cpu 2>&1 | tail -1 | tee -a $OUT/r02_g2b.log
timeout 300 python bench.py --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-600 | tee -a $OUT/r02_g2b.log
grep -v "^\s*$" $OUT/r02_xchg.log | grep -v "Traceback\|File \"/opt" | tail -30 | tee -a $OUT/r02_g2b.log
