#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== N=$N DataStream programs" | tee $OUT/r02_g18.log
timeout 1200 $TR --master-port 29613 tests/dist_nccl_check.py --more > $OUT/r02_nccl_check_g18.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_g18.log
grep -h "OK\|Error\|asof\|q2_q21" $OUT/r02_nccl_check_g18.log | tail -12 | tee -a $OUT/r02_g18.log
echo "== N=$N exchange check" | tee -a $OUT/r02_g18.log
timeout 600 $TR --master-port 29611 tests/dist_xchg_check.py > $OUT/r02_xchg_g18.log 2>&1; echo "xchg rc=$?" | tee -a $OUT/r02_g18.log
grep -h "OK\|Error" $OUT/r02_xchg_g18.log | tail -2 | tee -a $OUT/r02_g18.log
echo "== N=$N bench" | tee -a $OUT/r02_g18.log
SECONDS=0
timeout 1500 $TR --master-port 29622 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r02_bench_g18_n$N.json 2> $OUT/r02_bench_g18_n$N.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_g18.log
N=$N python - <<'PY' | tee -a gpurun_out/r02_g18.log
import json, os
d=json.loads(open(f"gpurun_out/r02_bench_g18_n{os.environ['N']}.json").read().strip().splitlines()[-1])
q=d['q3']
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('host_binding'))
print('q3 strong', q['seconds'], q.get('exchanges'), q.get('exchanges_via_peer_memory'), 'weak', q['weak']['seconds'], q['weak'].get('shuffle_bytes_over_nvlink'))
print('q5', d['q5']['seconds'], d['q5']['result'][:2], 'asof', d['asof']['seconds'], d['asof']['rows_per_s'], d['asof']['checksum'], d['asof'].get('join_kernels',{}).get('ms'))
PY
tail -2 $OUT/r02_bench_g18_n$N.err | cut -c1-300 | tee -a $OUT/r02_g18.log
echo done | tee -a $OUT/r02_g18.log
