#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== kernel tests (dyn plan col-col) + Q5 on one GPU" | tee $OUT/r02_g15.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "dynamic or dense" 2>&1 | tail -6 | tee -a $OUT/r02_g15.log
timeout 300 python bench.py --only-q5 --no-cpu 2>&1 | tail -1 | cut -c1-260 | tee -a $OUT/r02_g15.log
echo "== 2-rank pytest (test_gpu_multirank)" | tee -a $OUT/r02_g15.log
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 | tee -a $OUT/r02_g15.log
echo "== N=2 exchange + DataStream programs" | tee -a $OUT/r02_g15.log
timeout 600 $TR --master-port 29611 tests/dist_xchg_check.py > $OUT/r02_xchg_g15.log 2>&1; echo "xchg rc=$?" | tee -a $OUT/r02_g15.log
timeout 900 $TR --master-port 29613 tests/dist_nccl_check.py --more > $OUT/r02_nccl_check_g15.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_g15.log
grep -h "OK\|Error" $OUT/r02_xchg_g15.log $OUT/r02_nccl_check_g15.log | tail -4 | tee -a $OUT/r02_g15.log
echo "== N=2 bench" | tee -a $OUT/r02_g15.log
SECONDS=0
timeout 1500 $TR --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r02_bench_g15_n2.json 2> $OUT/r02_bench_g15_n2.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_g15.log
python - <<'PY' | tee -a gpurun_out/r02_g15.log
import json
d=json.loads(open('gpurun_out/r02_bench_g15_n2.json').read().strip().splitlines()[-1])
q=d['q3']
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('host_binding'))
print('q3 strong', q['seconds'], q.get('exchanges'), q.get('exchanges_via_peer_memory'), 'weak', q['weak']['seconds'])
print('q5', d['q5']['seconds'], d['q5']['result'][:2], 'asof', d['asof']['seconds'], d['asof']['checksum'], d['asof'].get('join_kernels',{}).get('ms'))
PY
echo "== reference arm under torchrun" | tee -a $OUT/r02_g15.log
timeout 600 $TR --master-port 29631 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200 | tee -a $OUT/r02_g15.log
echo done | tee -a $OUT/r02_g15.log
