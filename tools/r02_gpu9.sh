#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== tests (range node)" | tee $OUT/r02_g9.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_gpu_zz_more_api.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee -a $OUT/r02_g9.log
echo "== Q5 / Q3" | tee -a $OUT/r02_g9.log
timeout 300 python bench.py --only-q5 --no-cpu 2>&1 | tail -1 | cut -c1-300 | tee -a $OUT/r02_g9.log
timeout 300 python bench.py --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-300 | tee -a $OUT/r02_g9.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $OUT/r02_launches_q5_b.csv python bench.py --only-q5 --q3-steps 1 --no-cpu > $OUT/r02_q5_ncu_b.log 2>&1
echo "== ncu --set full: dynamic plan kernel on Q6's aggregate" | tee -a $OUT/r02_g9.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_agg_dyn -s 3 -c 1 -o $OUT/r02_prof_dyn python bench.py --steps 3 --no-e2e --no-q3 --no-cpu > /dev/null 2>&1
ls -la $OUT/r02_prof_dyn.ncu-rep | tee -a $OUT/r02_g9.log
echo done | tee -a $OUT/r02_g9.log
