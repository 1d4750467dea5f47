#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== dyn plan tests" | tee $OUT/r02_g10.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee -a $OUT/r02_g10.log
echo "== headline + q6 (dyn plan, grid sized to shared memory)" | tee -a $OUT/r02_g10.log
timeout 300 python bench.py --steps 5 --no-e2e --no-q3 --no-cpu 2>&1 | tail -1 > $OUT/r02_g10_bench.json
python - <<'PY' | tee -a gpurun_out/r02_g10.log
import json
d = json.loads(open("gpurun_out/r02_g10_bench.json").read())
print("value", d["value"], "frac", d["roofline"]["frac"])
print("q6", json.dumps(d.get("q6") or d.get("config", {}).get("q6") or {k: v for k, v in d.items() if "q6" in k})[:600])
PY
echo "== ncu --set full: dyn kernel after" | tee -a $OUT/r02_g10.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_agg_dyn -s 3 -c 1 -o $OUT/r02_prof_dyn_after python bench.py --steps 3 --no-e2e --no-q3 --no-cpu > /dev/null 2>&1
ls -la $OUT/r02_prof_dyn_after.ncu-rep | tee -a $OUT/r02_g10.log
echo done | tee -a $OUT/r02_g10.log
