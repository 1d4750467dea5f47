#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== full gpu suite on a 2-GPU box (multirank tests included)" | tee $OUT/r02_f4.log
SECONDS=0
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee -a $OUT/r02_f4.log
echo "suite wall=${SECONDS}s" | tee -a $OUT/r02_f4.log
echo "== N=2 bench" | tee -a $OUT/r02_f4.log
SECONDS=0
timeout 1500 $TR --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r02_bench_f4_n2.json 2> $OUT/r02_bench_f4_n2.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_f4.log
python - <<'PY' | tee -a gpurun_out/r02_f4.log
import json
d=json.loads(open('gpurun_out/r02_bench_f4_n2.json').read().strip().splitlines()[-1])
q=d['q3']
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('host_binding'))
print('q3 strong', q['seconds'], q.get('exchanges'), q.get('exchanges_via_peer_memory'), 'weak', q['weak']['seconds'])
print('q5', d['q5']['seconds'], d['q5']['result'][:2], 'asof', d['asof']['seconds'], d['asof']['rows_per_s'], d['asof']['checksum'])
PY
echo done | tee -a $OUT/r02_f4.log
