"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:  python tools/summarize_launches.py in.csv out.md "<command>" """
import csv
import re
import sys
from collections import defaultdict

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
h = rows[0]
ik, iv, ig = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = defaultdict(lambda: [0, 0.0, 0.0])
n = 0
for r in rows[1:]:
    us = float(r[iv].replace(",", "")) / 1000.0
    ctas = 1
    for v in re.findall(r"\d+", r[ig]):
        ctas *= int(v)
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()
    a = agg[name]; a[0] += 1; a[1] += us
    if ctas >= 148:
        a[2] += us
    n += 1
tot = sum(a[1] for a in agg.values()); big = sum(a[2] for a in agg.values())
with open(dst, "w") as f:
    f.write(f"# ncu launch list summary: `{cmd}` (NVTX range `timed`)\n\n")
    f.write(f"One timed step = one 18.2 M-parameter batch-256 proof, {n} launches, sum of kernel durations {tot / 1000:.2f} ms (serialised, cold cache;\n"
            f"the step itself is shorter because layers and sub-proofs overlap on streams); launches with >= 148 CTAs account for {big / 1000:.2f} ms.\n"
            "`ZKDL_PROVE_THREADS=0` issues every launch from the calling thread so that the NVTX range scopes them (the default feeds each stream\n"
            "from its own host thread; NVTX ranges are per thread).  Raw list: the `.csv` of the same name.\n\n")
    f.write("| kernel | launches | total µs | share | of which grids >= 148 CTAs µs |\n|---|---:|---:|---:|---:|\n")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{name}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[2]:.1f} |\n")
print(open(dst).read()[:3000])
