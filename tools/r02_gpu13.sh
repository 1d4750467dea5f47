#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== as-of kernel tests" | tee $OUT/r02_g13.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "asof" 2>&1 | tail -5 | tee -a $OUT/r02_g13.log
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_zz_more_api.py -m gpu -q -p no:cacheprovider -k "asof or ts or window" 2>&1 | tail -4 | tee -a $OUT/r02_g13.log
echo "== as-of bench (240 M rows and 1.26 B rows)" | tee -a $OUT/r02_g13.log
timeout 300 python bench.py --only-asof --asof-quotes 200000000 --no-cpu 2>&1 | tail -1 > $OUT/r02_g13_asof_small.json
timeout 300 python bench.py --only-asof --no-cpu 2>&1 | tail -1 > $OUT/r02_g13_asof.json
python - <<'PY' | tee -a gpurun_out/r02_g13.log
import json
for f in ("r02_g13_asof_small", "r02_g13_asof"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read())
        a = d["asof"]
        print(f, "seconds", a["seconds"], "rows/s", a["rows_per_s"], "checksum", a["checksum"], "n", a["trades_out"], "frac", a["roofline"]["frac"])
        print("   kernels", json.dumps(a.get("join_kernels"))[:300])
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.json").read()[-600:])
PY
echo "== Q3 / Q5 (raw stream handle)" | tee -a $OUT/r02_g13.log
timeout 300 python bench.py --only-q3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g13.log
timeout 300 python bench.py --only-q5 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/r02_g13.log
echo "== launch list, as-of 240 M rows" | tee -a $OUT/r02_g13.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $OUT/r02_launches_asof_cta2.csv python bench.py --only-asof --asof-quotes 200000000 --no-cpu > $OUT/r02_asof_ncu_cta2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_asof_sweep -s 2 -c 1 -o $OUT/r02_prof_asof_sweep_cta2 python bench.py --only-asof --asof-quotes 200000000 --no-cpu > /dev/null 2>&1
ls -la $OUT/r02_prof_asof_sweep_cta2.ncu-rep | tee -a $OUT/r02_g13.log
echo done | tee -a $OUT/r02_g13.log
