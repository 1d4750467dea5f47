#!/usr/bin/env bash
# Round 2, single-GPU validation of the final tree: the whole -m gpu suite, smoke(), the default bench line, launch lists.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== full GPU suite" | tee $OUT/r02_f1.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 | tee -a $OUT/r02_f1.log
echo "== smoke" | tee -a $OUT/r02_f1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $OUT/r02_f1.log
echo "== default bench line" | tee -a $OUT/r02_f1.log
/usr/bin/time -v timeout 1500 python bench.py > $OUT/r02_bench_final_n1.json 2> $OUT/r02_bench_final_n1.err; echo "rc=$?" | tee -a $OUT/r02_f1.log
grep "Elapsed (wall" $OUT/r02_bench_final_n1.err | tee -a $OUT/r02_f1.log
python - <<'PY' | tee -a $OUT/r02_f1.log
import json
d=json.loads(open('gpurun_out/r02_bench_final_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['value'])
print('q6', d.get('q6'))
for k in ('q3','q5','asof'):
    x=d[k]; print(k, x.get('seconds'), x.get('rows_per_s'), (x.get('roofline') or {}).get('frac'), x.get('error'))
print('parquet', {k:(v.get('ms') if isinstance(v,dict) else v) for k,v in d['e2e_parquet'].items() if k.startswith(('host','device'))})
print('cpu', d['cpu_baseline']['value'], d['clocks'])
PY
echo "== reference arm" | tee -a $OUT/r02_f1.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-300 | tee -a $OUT/r02_f1.log
echo "== launch lists: partition / hash aggregate microbench, as-of" | tee -a $OUT/r02_f1.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $OUT/r02_launches_partition.csv python tools/prof_partition.py > $OUT/r02_partition_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_asof_e.csv python bench.py --only-asof --no-cpu --asof-quotes 200000000 > $OUT/r02_asof_ncu_e.log 2>&1
timeout 300 python bench.py --only-asof --no-cpu --asof-quotes 200000000 2>&1 | tail -1 | cut -c1-400 | tee -a $OUT/r02_f1.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $OUT/r02_launches_q5.csv python bench.py --only-q5 --q3-steps 1 --no-cpu > $OUT/r02_q5_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $OUT/r02_launches_q3_e.csv python bench.py --only-q3 --q3-steps 1 --no-cpu > $OUT/r02_q3_ncu_e.log 2>&1
timeout 300 python bench.py --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_f1.log
echo done | tee -a $OUT/r02_f1.log
