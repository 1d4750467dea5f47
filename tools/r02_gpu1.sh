#!/usr/bin/env bash
# Round 2, GPU call 1: the whole -m gpu suite (no -x), the Parquet bench leg, launch lists of as-of / Q3 / Q5.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== 1. full GPU suite" | tee $OUT/r02_g1.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 | tee $OUT/r02_pytest1.log | tail -15 | tee -a $OUT/r02_g1.log
echo "== 2. Parquet bench leg" | tee -a $OUT/r02_g1.log
timeout 600 python bench.py --only-parquet --parquet-sf 10 > $OUT/r02_parquet_bench.json 2> $OUT/r02_parquet_bench.err; tail -c 3000 $OUT/r02_parquet_bench.json | tee -a $OUT/r02_g1.log; tail -5 $OUT/r02_parquet_bench.err | tee -a $OUT/r02_g1.log
echo "== 3. launch lists" | tee -a $OUT/r02_g1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_asof.csv python bench.py --only-asof --no-cpu > $OUT/r02_asof_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $OUT/r02_launches_q3.csv python bench.py --only-q3 --q3-steps 1 --no-cpu > $OUT/r02_q3_ncu.log 2>&1
echo "== 4. bench timings (asof / q3+q5, QK_PROFILE)" | tee -a $OUT/r02_g1.log
timeout 600 python bench.py --only-asof --no-cpu 2>&1 | tail -2 | tee $OUT/r02_asof.json | cut -c1-1500 | tee -a $OUT/r02_g1.log
QK_PROFILE=1 timeout 600 python bench.py --only-q3 --no-cpu 2>&1 | tail -2 | tee $OUT/r02_q3_profile.json | cut -c1-3000 | tee -a $OUT/r02_g1.log
echo done | tee -a $OUT/r02_g1.log
