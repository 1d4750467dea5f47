#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== dyn plan test" | tee $OUT/r02_g17.log
timeout 600 python -m pytest "tests/test_gpu_kernels.py::test_dense_agg_dynamic_fused_plan" -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee -a $OUT/r02_g17.log
echo "== N=2 DataStream programs (time-range as-of, all 22 TPC-H programs)" | tee -a $OUT/r02_g17.log
timeout 1200 $TR --master-port 29613 tests/dist_nccl_check.py --more > $OUT/r02_nccl_check_g17.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_g17.log
grep -h "OK\|Error\|asof" $OUT/r02_nccl_check_g17.log | tail -12 | tee -a $OUT/r02_g17.log
echo "== as-of N=2 profile" | tee -a $OUT/r02_g17.log
QK_PROFILE=1 timeout 900 $TR --master-port 29641 bench.py --gpus 2 --only-asof --no-cpu > $OUT/r02_g17_asof_prof.json 2> $OUT/r02_g17_asof_prof.err; echo "rc=$?" | tee -a $OUT/r02_g17.log
grep "asof profile_ms" $OUT/r02_g17_asof_prof.err | tail -1 | cut -c1-1200 | tee -a $OUT/r02_g17.log
echo "== N=2 bench" | tee -a $OUT/r02_g17.log
SECONDS=0
timeout 1500 $TR --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r02_bench_g17_n2.json 2> $OUT/r02_bench_g17_n2.err; echo "rc=$? wall=${SECONDS}s" | tee -a $OUT/r02_g17.log
python - <<'PY' | tee -a gpurun_out/r02_g17.log
import json
d=json.loads(open('gpurun_out/r02_bench_g17_n2.json').read().strip().splitlines()[-1])
q=d['q3']
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d.get('host_binding'))
print('q3 strong', q['seconds'], q.get('exchanges'), q.get('exchanges_via_peer_memory'), 'weak', q['weak']['seconds'])
print('q5', d['q5']['seconds'], d['q5']['result'][:2], 'asof', d['asof']['seconds'], d['asof']['rows_per_s'], d['asof']['checksum'], d['asof'].get('join_kernels',{}).get('ms'))
PY
tail -3 $OUT/r02_bench_g17_n2.err | cut -c1-300 | tee -a $OUT/r02_g17.log
echo done | tee -a $OUT/r02_g17.log
