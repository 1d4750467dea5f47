#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== as-of kernel tests (1 GPU)" | tee $OUT/r02_g19.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "asof" 2>&1 | tail -3 | tee -a $OUT/r02_g19.log
echo "== N=2 DataStream programs" | tee -a $OUT/r02_g19.log
timeout 1200 $TR --master-port 29613 tests/dist_nccl_check.py --more > $OUT/r02_nccl_check_g19.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_g19.log
grep -h "OK\|Error\|asof" $OUT/r02_nccl_check_g19.log | tail -8 | tee -a $OUT/r02_g19.log
echo "== as-of N=2" | tee -a $OUT/r02_g19.log
timeout 900 $TR --master-port 29642 bench.py --gpus 2 --only-asof --no-cpu 2>/dev/null | tail -1 | cut -c1-420 | tee -a $OUT/r02_g19.log
QK_PROFILE=1 timeout 900 $TR --master-port 29641 bench.py --gpus 2 --only-asof --no-cpu > /dev/null 2> $OUT/r02_g19_asof_prof.err
grep "asof profile_ms" $OUT/r02_g19_asof_prof.err | tail -1 | cut -c1-700 | tee -a $OUT/r02_g19.log
echo done | tee -a $OUT/r02_g19.log
