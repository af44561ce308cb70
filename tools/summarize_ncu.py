#!/usr/bin/env python
"""Turn gpurun_out ncu artefacts into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
    python tools/summarize_ncu.py raw gpurun_out/prof_gemm2.ncu-rep profiles/r01_gemm_ncu.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                "(cold-cache, serialised: compare SHARES, not absolutes)\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {v[1] / tot:.3f} |\n")
    print(open(dst).read())


def raw(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({rep})\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[idx['Kernel Name']][:110]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
