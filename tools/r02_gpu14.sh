#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== full gpu suite" | tee $OUT/r02_g14.log
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 | tee -a $OUT/r02_g14.log
echo "suite wall=${SECONDS}s" | tee -a $OUT/r02_g14.log
echo "== smoke" | tee -a $OUT/r02_g14.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3 | tee -a $OUT/r02_g14.log
echo "== default bench line, N=1" | tee -a $OUT/r02_g14.log
SECONDS=0
timeout 1500 python bench.py > $OUT/r02_bench_g14_n1.json 2> $OUT/r02_bench_g14_n1.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_g14.log
python - <<'PY' | tee -a gpurun_out/r02_g14.log
import json
d=json.loads(open('gpurun_out/r02_bench_g14_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['value'])
print('q6', {k: d['q6'].get(k) for k in ('kernel', 'ms', 'error')}, d['q6'].get('roofline', {}).get('frac'))
for k in ('q3','q5','asof'):
    x=d[k]; print(k, x.get('seconds'), x.get('rows_per_s'), (x.get('roofline') or {}).get('frac'), x.get('error'))
print('asof kernels', json.dumps(d['asof'].get('join_kernels'))[:200])
print('parquet', {k:(v.get('ms') if isinstance(v,dict) else v) for k,v in d['e2e_parquet'].items() if k.startswith(('host','device'))})
print('cpu', d['cpu_baseline']['value'], d['clocks'])
PY
echo "== reference arm" | tee -a $OUT/r02_g14.log
SECONDS=0
timeout 900 python bench.py --impl reference > $OUT/r02_bench_g14_ref.json 2> $OUT/r02_bench_g14_ref.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_g14.log
tail -1 $OUT/r02_bench_g14_ref.json | cut -c1-400 | tee -a $OUT/r02_g14.log
echo done | tee -a $OUT/r02_g14.log
