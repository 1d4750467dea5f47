#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== full gpu suite" | tee $OUT/r02_f3.log
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee -a $OUT/r02_f3.log
echo "suite wall=${SECONDS}s" | tee -a $OUT/r02_f3.log
echo "== smoke" | tee -a $OUT/r02_f3.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 | tee -a $OUT/r02_f3.log
echo "== default bench line, N=1" | tee -a $OUT/r02_f3.log
SECONDS=0
timeout 1500 python bench.py > $OUT/r02_bench_f3_n1.json 2> $OUT/r02_bench_f3_n1.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_f3.log
python - <<'PY' | tee -a gpurun_out/r02_f3.log
import json
d=json.loads(open('gpurun_out/r02_bench_f3_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['value'], 'launches', d.get('gpu_launches'))
print('q6', {k: d['q6'].get(k) for k in ('kernel', 'ms', 'error')}, d['q6'].get('roofline', {}).get('frac'))
for k in ('q3','q5','asof'):
    x=d[k]; print(k, x.get('seconds'), x.get('rows_per_s'), (x.get('roofline') or {}).get('frac'), x.get('error'))
print('asof kernels', json.dumps(d['asof'].get('join_kernels'))[:200])
print('parquet', {k:(v.get('ms') if isinstance(v,dict) else v) for k,v in d['e2e_parquet'].items() if k.startswith(('host','device'))})
print('cpu', d['cpu_baseline']['value'], d['clocks'])
PY
echo "== launch list Q5" | tee -a $OUT/r02_f3.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $OUT/r02_launches_q5_c.csv python bench.py --only-q5 --q3-steps 1 --no-cpu > $OUT/r02_q5_ncu_c.log 2>&1
echo done | tee -a $OUT/r02_f3.log
