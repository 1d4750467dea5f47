#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== dyn plan tests" | tee $OUT/r02_g11.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 | tee -a $OUT/r02_g11.log
echo "== headline + q6 + shape sweep" | tee -a $OUT/r02_g11.log
timeout 300 python bench.py --steps 5 --no-e2e --no-q3 --no-cpu --dyn-sweep 2>&1 | tail -1 > $OUT/r02_g11_bench.json
python - <<'PY' | tee -a gpurun_out/r02_g11.log
import json
d = json.loads(open("gpurun_out/r02_g11_bench.json").read())
print("value", d["value"], "frac", d["roofline"]["frac"], "parity", d.get("parity") or d["config"].get("parity"))
q6 = d["q6"]
print("q6", q6.get("kernel"), q6.get("ms"), q6.get("roofline", {}).get("frac"), q6.get("rows_passing"), q6.get("rows_passing_torch"), q6.get("rel_err_vs_torch_fp64"), q6.get("error"))
for k, v in (q6.get("shape_sweep_ms") or {}).items(): print("  ", k, v)
PY
echo "== ncu --set full: dyn kernel, default shape" | tee -a $OUT/r02_g11.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_agg_dyn -s 3 -c 1 -o $OUT/r02_prof_dyn_tile python bench.py --steps 3 --no-e2e --no-q3 --no-cpu > /dev/null 2>&1
ls -la $OUT/r02_prof_dyn_tile.ncu-rep | tee -a $OUT/r02_g11.log
echo done | tee -a $OUT/r02_g11.log
