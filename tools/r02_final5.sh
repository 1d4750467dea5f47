#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== N=2 DataStream programs after the silent-stream change" | tee $OUT/r02_f5.log
timeout 100 $TR --master-port 29613 tests/dist_nccl_check.py > $OUT/r02_nccl_check_f5.log 2>&1; echo "nccl_check rc=$?" | tee -a $OUT/r02_f5.log
grep -h "OK\|Error\|case_q3\|case_q5" $OUT/r02_nccl_check_f5.log | tail -5 | tee -a $OUT/r02_f5.log
echo "== Q3 / Q5 at N=2" | tee -a $OUT/r02_f5.log
timeout 60 $TR --master-port 29622 bench.py --gpus 2 --only-q3 --no-cpu 2>/dev/null | tail -1 > $OUT/r02_f5_q3.json
timeout 60 $TR --master-port 29623 bench.py --gpus 2 --only-q5 --no-cpu 2>/dev/null | tail -1 > $OUT/r02_f5_q5.json
python - <<'PY' | tee -a gpurun_out/r02_f5.log
import json
try:
    q = json.loads(open('gpurun_out/r02_f5_q3.json').read())['q3']
    print('q3 strong', q['seconds'], q.get('exchanges'), q.get('exchanges_via_peer_memory'), 'weak', q['weak']['seconds'], 'top', str(q.get('top1') or q.get('result', ''))[:80])
except Exception as e: print('q3 failed', e)
try:
    q = json.loads(open('gpurun_out/r02_f5_q5.json').read())['q5']
    print('q5', q['seconds'], q['result'][:2])
except Exception as e: print('q5 failed', e)
PY
echo done | tee -a $OUT/r02_f5.log
