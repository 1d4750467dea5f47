#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== tests" | tee $OUT/r02_g7.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_gpu_zz_more_api.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee -a $OUT/r02_g7.log
echo "== Q3 N=1" | tee -a $OUT/r02_g7.log
timeout 300 python bench.py --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g7.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 900 --csv --log-file $OUT/r02_launches_q3_d.csv python bench.py --only-q3 --q3-steps 1 --no-cpu > $OUT/r02_q3_ncu_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compact_mask -s 4 -c 2 -o $OUT/r02_prof_mask2 python bench.py --only-q3 --q3-steps 1 --no-cpu > /dev/null 2>&1
echo "== Q3 N=2, shuffle all vs replicate small builds" | tee -a $OUT/r02_g7.log
timeout 300 $TR --master-port 29620 bench.py --gpus 2 --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g7.log
timeout 300 $TR --master-port 29621 bench.py --gpus 2 --only-q3 --no-cpu --replicate-builds 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g7.log
echo "== full default bench N=2" | tee -a $OUT/r02_g7.log
timeout 1500 $TR --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r02_bench_n2.json 2> $OUT/r02_bench_n2.err; echo "rc=$?" | tee -a $OUT/r02_g7.log
tail -c 3000 $OUT/r02_bench_n2.json | tee -a $OUT/r02_g7.log; grep -v "^\*\|OMP_NUM\|^$" $OUT/r02_bench_n2.err | tail -8 | tee -a $OUT/r02_g7.log
echo done | tee -a $OUT/r02_g7.log
