#!/usr/bin/env python
"""
Aggregate the stall samples of an ncu --set full capture per SOURCE LINE of the calling kernel (inlined helpers are
attributed to the line of the kernel that called them).

    python tools/ncu_lines.py <report.ncu-rep> <kernel-name-substring> <libfoldingdiff_b200.so> <source.cuh> [top] [section]
    (section: index of the kernel section of a multi-kernel report, default 0)

ncu's CSV source page carries SASS only; the line table comes from `nvdisasm -gi` of the cubin inside the shipped .so
(same build as the capture: instruction k of the kernel in both listings is the same instruction).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kern, lib, src_file = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    rows = rows[heads[sec]:heads[sec + 1] if sec + 1 < len(heads) else len(rows)]
    name = rows[0][1]
    assert kern in name, name
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", kern)
    starts = [i for i, l in enumerate(sass) if l.startswith("//--------------------- .text.") and mangled_hint in l]
    best = None
    for st in starts:  # pick the function whose instruction count matches the capture
        en = next(i for i in range(st + 1, len(sass)) if sass[i].startswith("//--------------------- ") or i == len(sass) - 1)
        n = sum(1 for l in sass[st:en] if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l))
        if n == len(data):
            best = (st, en)
    assert best, "no function with %d instructions" % len(data)
    cur, table = None, []
    for l in sass[best[0]:best[1]]:
        if "//## File" in l:
            files = re.findall(r'"([^"]+)", line (\d+)', l)
            cur = files  # innermost first, outermost (the kernel's own line) last
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            table.append(cur)
    base = os.path.basename(src_file)
    src = open(src_file).read().split("\n")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r, loc in zip(data, table):
        line = 0
        for f, n in (loc or []):
            if os.path.basename(f) == base:
                line = int(n)  # keeps the OUTERMOST occurrence in the kernel's file
        a = agg[line]
        a[0] += int(r[ix["# Samples"]] or 0)
        a[1] += int(r[ix["Instructions Executed"]] or 0)
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                a[2][c.replace("stall_", "")] += v
    tot = sum(a[0] for a in agg.values()) or 1
    print(f"# {name[:90]}\n# {tot} samples, {sum(a[1] for a in agg.values())} warp instructions")
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        txt = src[line - 1].strip()[:95] if 0 < line <= len(src) else ""
        print(f"{a[0]:5d} {100 * a[0] / tot:5.1f}%  inst {a[1]:8d}  L{line:<4d} {dict(a[2].most_common(2))}  {txt}")


if __name__ == "__main__":
    main()
