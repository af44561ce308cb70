#!/usr/bin/env python
"""Turn gpurun_out ncu artefacts into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
    python tools/summarize_ncu.py raw gpurun_out/prof_gemm2.ncu-rep profiles/r01_gemm_ncu.md
    python tools/summarize_ncu.py step gpurun_out/step_metrics.csv profiles/r02_step_ncu.json
        one reverse step (tools/gpu_ncu_step.sh: --metrics time, DRAM bytes, tensor-pipe, issue) -> per-category JSON that
        bench.py reads for `roofline.traffic` / tensor-pipe %.  The JSON records a hash of the CUDA sources it was
        captured from (`source_sha`, computed HERE from the tree the capture ran on - run this right after the capture
        returns); bench.py recomputes the hash and flags the figures as stale when the kernels have changed since.
"""
import collections
import csv
import os
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


def source_sha():
    """sha256 over the CUDA sources of the library (what bench.py recomputes to detect a stale capture)."""
    import hashlib
    import os
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "foldingdiff_b200", "csrc")
    h = hashlib.sha256()
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".hpp")):
            h.update(name.encode())
            h.update(open(os.path.join(csrc, name), "rb").read())
    return h.hexdigest()[:16]


def step_category(name, n_gemm_seen, layers=12):
    """Launch order of one tensor-core reverse step: embed, L x {gemm qkv, attention, gemm out, gemm ffn1, gemm ffn2},
    gemm head, tail (api.cu: run_encoder)."""
    if "embed_kernel" in name:
        return "embed"
    if "tail_kernel" in name:
        return "tail_posterior"
    if "attention" in name:
        return "attention"
    if "layernorm" in name:
        return "layernorm"
    if "gemm" in name:
        if n_gemm_seen >= 4 * layers:
            return "gemm_head"
        return ["gemm_qkv", "gemm_attn_out", "gemm_ffn1", "gemm_ffn2"][n_gemm_seen % 4]
    return "other"


def step(src, dst):
    import json
    import subprocess as sp
    lines = [l for l in open(src) if not l.startswith("==")]
    per_launch = collections.OrderedDict()  # ID -> {name, metric: value}
    for row in csv.DictReader(lines):
        d = per_launch.setdefault(row["ID"], {"name": row["Kernel Name"]})
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[row["Metric Name"]] = v * scale
    cats = collections.OrderedDict()
    gemms = 0
    for d in per_launch.values():
        c = step_category(d["name"], gemms)
        if "gemm" in d["name"]:
            gemms += 1
        a = cats.setdefault(c, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "tensor_pct_x_time": 0.0,
                                "issue_pct_x_time": 0.0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["time_us"] += t
        a["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor_pct_x_time"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["issue_pct_x_time"] += t * d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
    sha_file = os.path.join(os.path.dirname(os.path.abspath(src)), "step_source_sha.txt")  # written on the GPU box by the capture
    sha = open(sha_file).read().strip() if os.path.isfile(sha_file) else source_sha()
    out = {"source_sha": sha, "captured_from": src, "launches": len(per_launch),
           "head": sp.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
           "note": "ncu --clock-control none, one reverse step of BASELINE config 2 (512 chains, len 50-127), cold-cache serialised launches: "
                   "per-launch times are NOT bench values (compare shares); DRAM bytes and pipe percentages are what this file is for",
           "kernels": {}}
    tot_t = sum(a["time_us"] for a in cats.values()) or 1.0
    for c, a in cats.items():
        t = a["time_us"] or 1.0
        out["kernels"][c] = {"launches": a["launches"], "time_us": round(a["time_us"], 1), "share": round(a["time_us"] / tot_t, 4),
                             "dram_bytes_per_launch": round((a["dram_read_bytes"] + a["dram_write_bytes"]) / a["launches"]),
                             "dram_read_bytes": round(a["dram_read_bytes"]), "dram_write_bytes": round(a["dram_write_bytes"]),
                             "tensor_pipe_active_pct": round(a["tensor_pct_x_time"] / t, 2),
                             "issue_active_pct": round(a["issue_pct_x_time"] / t, 2)}
    out["dram_bytes_per_reverse_step"] = round(sum(a["dram_read_bytes"] + a["dram_write_bytes"] for a in cats.values()))
    out["tensor_pipe_active_pct_time_weighted"] = round(sum(a["tensor_pct_x_time"] for a in cats.values()) / tot_t, 2)
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "step": step}[sys.argv[1]](sys.argv[2], sys.argv[3])
