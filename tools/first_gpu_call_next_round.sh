#!/usr/bin/env bash
# One gpurun call that validates and measures everything written after round 1's last GPU session.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call_next_round.sh'
# Outputs land in gpurun_out/ (merged back by gpurun).  Each step is bounded by its own timeout so a hang cannot
# take the box to gpurun's limit.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== 1. tests collected last (never run on hardware before)" | tee $OUT/r02_first.log
timeout 900 python -m pytest tests/test_gpu_zz_more_api.py tests/test_gpu_zz_parquet.py -m gpu -q 2>&1 | tail -25 | tee -a $OUT/r02_first.log
echo "== 2. Q1 from Parquet files: Arrow host reader vs device decode (none / snappy / zstd)" | tee -a $OUT/r02_first.log
timeout 600 python bench.py --only-parquet --parquet-sf 10 2>&1 | tail -3 | tee -a $OUT/r02_first.log
echo "== 3. launch list + one full capture of the decode / inflate kernels (SF-1 file, three codecs)" | tee -a $OUT/r02_first.log
cat > /tmp/pq_prof.py <<'PY'
import sys; sys.path.insert(0, "tests")
import pyarrow.parquet as pq, torch
from oracle import tpch_gen as G
import parquet_cases as P
names = ["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"]
t = G.to_arrow(G.gen_lineitem(1, columns=names))
for codec in (None, "snappy", "zstd"):
    path = f"/tmp/prof_{codec}.parquet"
    pq.write_table(t, path, compression=codec, row_group_size=100_000)
    for _ in range(2):
        d = P.read(path, torch.device("cuda", 0), names)
    torch.cuda.synchronize()
    print(codec, len(d))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/r02_parquet_launches.csv python /tmp/pq_prof.py >> $OUT/r02_first.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pq_ -c 6 -o $OUT/r02_parquet_kernels python /tmp/pq_prof.py >> $OUT/r02_first.log 2>&1
echo "== 4. default bench line (regression check of the validated path)" | tee -a $OUT/r02_first.log
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $OUT/r02_bench_n1.json | cut -c1-400 | tee -a $OUT/r02_first.log
echo "done" | tee -a $OUT/r02_first.log
