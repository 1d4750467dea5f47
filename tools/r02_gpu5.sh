#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== Q3 N=2 QK_PROFILE" | tee $OUT/r02_g5.log
QK_PROFILE=1 timeout 300 $TR --master-port 29620 bench.py --gpus 2 --only-q3 --no-cpu 2>&1 | tail -1 | tee $OUT/r02_q3_n2_profile.json | cut -c1-2600 | tee -a $OUT/r02_g5.log
echo "== Q3 N=1 QK_PROFILE" | tee -a $OUT/r02_g5.log
QK_PROFILE=1 timeout 300 python bench.py --only-q3 --no-cpu 2>&1 | tail -1 | tee $OUT/r02_q3_n1_profile.json | cut -c1-2600 | tee -a $OUT/r02_g5.log
echo "== new kernels: dyn plan + multirank tests" | tee -a $OUT/r02_g5.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider 2>&1 | tail -12 | tee -a $OUT/r02_g5.log
echo "== launch lists (new kernels)" | tee -a $OUT/r02_g5.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $OUT/r02_launches_q3_b.csv python bench.py --only-q3 --q3-steps 1 --no-cpu > $OUT/r02_q3_ncu_b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_asof_b.csv python bench.py --only-asof --no-cpu --asof-quotes 200000000 > $OUT/r02_asof_ncu_b.log 2>&1
echo "== Q1-shaped aggregates through the dynamic plan at SF-100 (variant 7) vs typed (3) vs interpreter (1)" | tee -a $OUT/r02_g5.log
for v in 3 7 1; do timeout 300 python bench.py --variant $v --steps 10 --no-e2e --no-q3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['kernel'], d['ms_per_step'], d['roofline']['frac'], d['parity'])" | tee -a $OUT/r02_g5.log; done
echo done | tee -a $OUT/r02_g5.log
