#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== failing dyn test, full traceback" | tee $OUT/r02_g16.log
timeout 600 python -m pytest "tests/test_gpu_kernels.py::test_dense_agg_dynamic_fused_plan" -m gpu -q -p no:cacheprovider -x 2>&1 | tail -45 > $OUT/r02_g16_dyn.log
tail -12 $OUT/r02_g16_dyn.log | tee -a $OUT/r02_g16.log
echo "== as-of N=2 with per-step profile (synchronising: times are serialised)" | tee -a $OUT/r02_g16.log
QK_PROFILE=1 timeout 900 $TR --master-port 29641 bench.py --gpus 2 --only-asof --no-cpu > $OUT/r02_g16_asof_n2.json 2> $OUT/r02_g16_asof_n2.err; echo "rc=$?" | tee -a $OUT/r02_g16.log
grep "asof profile_ms" $OUT/r02_g16_asof_n2.err | tail -1 | cut -c1-1500 | tee -a $OUT/r02_g16.log
tail -1 $OUT/r02_g16_asof_n2.json | cut -c1-400 | tee -a $OUT/r02_g16.log
echo done | tee -a $OUT/r02_g16.log
