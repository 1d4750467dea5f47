#!/usr/bin/env python
"""Turns an ncu report / launch list brought back in gpurun_out/ into the small text summaries committed
under profiles/ (the .ncu-rep files themselves are scratch).

  python profiles/summarize.py rep gpurun_out/prof_q1.ncu-rep  > profiles/r01_q1_fused_tma.txt
  python profiles/summarize.py launches gpurun_out/launches.csv > profiles/r01_launches_q1.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
STALLS = "smsp__average_warps_issue_stalled_"


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        print(f"kernel: {d.get('Kernel Name', '?')}")
        for k in KEYS:
            if k in d:
                print(f"  {k:72s} {d[k]:>16s} {units[hdr.index(k)]}")
        st = sorted(((float(v or 0), k[len(STALLS):].replace('_per_issue_active.ratio', '')) for k, v in d.items()
                     if k.startswith(STALLS) and k.endswith("per_issue_active.ratio")), reverse=True)
        print("  top stall reasons (warps stalled per issued instruction):")
        for v, k in st[:6]:
            print(f"    {k:30s} {v:8.3f}")
        try:
            t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units[hdr.index("gpu__time_duration.sum")], 1e-9)
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * scale.get(units[hdr.index("dram__bytes_read.sum")], 1)
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * scale.get(units[hdr.index("dram__bytes_write.sum")], 1)
            print(f"  => DRAM traffic {(rd + wr) / 1e9:.3f} GB in {t * 1e3:.3f} ms = {(rd + wr) / t / 1e9:.0f} GB/s (cold-cache, under the profiler)")
        except Exception as e:  # noqa
            print("  (could not derive GB/s:", e, ")")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0].replace("qk::<unnamed>::", "")[:70]
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(k, [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"{len(rows) - 1} launches, {tot / 1e6:.3f} ms of device time in total (ncu: cold cache, serialised)")
    for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{v / 1e6:10.3f} ms {100 * v / tot:6.2f} % {c:5d}x  {k}")


def launches_dram(path, skip="k_synth"):
    """Launch list taken with gpu__time_duration.sum + dram__bytes_{read,write}.sum: per-kernel totals and the launches of the
    LAST repetition of the workload one by one (the first repetitions are warm-up)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per, order = {}, []
    for r in rows:
        i = r["ID"]
        if i not in per:
            per[i] = {"k": r["Kernel Name"].split("(")[0].replace("qk::<unnamed>::", "")[:48], "grid": r["Grid Size"]}
            order.append(i)
        per[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    agg = collections.OrderedDict()
    for i in order:
        d = per[i]
        a = agg.setdefault(d["k"], [0.0, 0.0, 0.0, 0])
        a[0] += d.get("gpu__time_duration.sum", 0.0); a[1] += d.get("dram__bytes_read.sum", 0.0); a[2] += d.get("dram__bytes_write.sum", 0.0); a[3] += 1
    tot = sum(v[0] for k, v in agg.items() if skip not in k)
    print(f"{len(order)} launches; device time without the data generator ({skip}): {tot / 1e6:.3f} ms (ncu: cold cache, serialised)")
    for k, (t, rd, wr, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if skip in k:
            continue
        print(f"{t / 1e6:9.3f} ms {100 * t / tot:6.2f} % {c:4d}x  DRAM rd {rd / 1e9:7.2f} GB wr {wr / 1e9:6.2f} GB  {(rd + wr) / max(t, 1) :7.1f} GB/s  {k}")
    last = [i for i in order if skip not in per[i]["k"]]
    last = last[len(last) // 2:]
    print("\n-- launches of the last repetition --")
    for i in last:
        d = per[i]
        t = d.get("gpu__time_duration.sum", 0.0)
        if t < 20e3:
            continue
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        print(f"{t / 1e3:9.1f} us  rd {rd / 1e6:8.1f} MB  wr {wr / 1e6:8.1f} MB  {(rd + wr) / max(t, 1):7.1f} GB/s  grid {d['grid']:>12s}  {d['k']}")


if __name__ == "__main__":
    {"rep": rep, "launches": launches, "launches_dram": launches_dram}[sys.argv[1]](sys.argv[2])
