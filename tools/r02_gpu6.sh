#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== tests" | tee $OUT/r02_g6.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee -a $OUT/r02_g6.log
echo "== Q3 / asof" | tee -a $OUT/r02_g6.log
timeout 300 python bench.py --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-400 | tee -a $OUT/r02_g6.log
timeout 300 python bench.py --only-asof --no-cpu --asof-quotes 200000000 2>&1 | tail -1 | cut -c1-500 | tee -a $OUT/r02_g6.log
echo "== dyn plan at SF-100 (Q1 shape, variant 7)" | tee -a $OUT/r02_g6.log
timeout 300 python bench.py --variant 7 --steps 10 --no-e2e --no-q3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['kernel'], d['ms_per_step'], d['roofline']['frac'], d['parity'])" | tee -a $OUT/r02_g6.log
echo "== launch lists" | tee -a $OUT/r02_g6.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $OUT/r02_launches_q3_c.csv python bench.py --only-q3 --q3-steps 1 --no-cpu > $OUT/r02_q3_ncu_c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_asof_c.csv python bench.py --only-asof --no-cpu --asof-quotes 200000000 > $OUT/r02_asof_ncu_c.log 2>&1
echo "== full default bench line" | tee -a $OUT/r02_g6.log
timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err; echo "rc=$?" | tee -a $OUT/r02_g6.log; tail -c 6000 $OUT/r02_bench_n1.json | tee -a $OUT/r02_g6.log; tail -5 $OUT/r02_bench_n1.err | tee -a $OUT/r02_g6.log
echo "== ncu --set full: mask, sweep" | tee -a $OUT/r02_g6.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compact_mask -s 4 -c 2 -o $OUT/r02_prof_mask python bench.py --only-q3 --q3-steps 1 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_asof_sweep -c 1 -o $OUT/r02_prof_sweep python bench.py --only-asof --no-cpu --asof-quotes 200000000 > /dev/null 2>&1
ls -la $OUT/*.ncu-rep | tail -3 | tee -a $OUT/r02_g6.log
echo done | tee -a $OUT/r02_g6.log
