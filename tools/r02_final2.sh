#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== default bench line, N=1" | tee $OUT/r02_f2.log
SECONDS=0
timeout 1500 python bench.py > $OUT/r02_bench_final_n1.json 2> $OUT/r02_bench_final_n1.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_f2.log
python - <<'PY' | tee -a $OUT/r02_f2.log
import json
d=json.loads(open('gpurun_out/r02_bench_final_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['value'])
print('q6', d.get('q6'))
for k in ('q3','q5','asof'):
    x=d[k]; print(k, x.get('seconds'), x.get('rows_per_s'), (x.get('roofline') or {}).get('frac'), x.get('error'))
print('parquet', {k:(v.get('ms') if isinstance(v,dict) else v) for k,v in d['e2e_parquet'].items() if k.startswith(('host','device'))})
print('cpu', d['cpu_baseline']['value'], d['clocks'])
PY
echo "== N=2: DataStream programs, full bench" | tee -a $OUT/r02_f2.log
timeout 600 $TR --master-port 29613 tests/dist_nccl_check.py --more > $OUT/r02_nccl_check_f.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_f2.log
grep -h "OK\|Error" $OUT/r02_nccl_check_f.log | tail -3 | tee -a $OUT/r02_f2.log
SECONDS=0
timeout 1500 $TR --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r02_bench_final_n2.json 2> $OUT/r02_bench_final_n2.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_f2.log
python - <<'PY' | tee -a $OUT/r02_f2.log
import json
d=json.loads(open('gpurun_out/r02_bench_final_n2.json').read().strip().splitlines()[-1])
q=d['q3']
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('host_binding'))
print('q3 strong', q['seconds'], q['exchanges'], q['exchanges_via_peer_memory'], 'weak', q['weak']['seconds'])
print('q5', d['q5']['seconds'], d['q5']['result'][:2], 'asof', d['asof']['seconds'], d['asof']['checksum'])
PY
grep -v "^\*\|OMP_NUM\|^$" $OUT/r02_bench_final_n2.err | tail -5 | tee -a $OUT/r02_f2.log
echo done | tee -a $OUT/r02_f2.log
