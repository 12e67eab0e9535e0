#!/usr/bin/env python
"""Stall samples and executed instructions per source line of an --import-source ncu report:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python profiles/tools/ncu_source_by_line.py src.csv [top_n]
Prints the per-file totals, the top lines, and for fused.cu the totals per kernel phase."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr = None, None
lines = []   # (file, line, src, samples, insts, stall dict)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None or cur is None:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    if len(r) > len(hdr):   # the source text contained the separator: re-join it
        extra = len(r) - len(hdr)
        r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
    d = dict(zip(hdr[4:], r[4:]))
    num = lambda v: int(v) if v and v.lstrip("-").isdigit() else 0
    samples = num(d.get("# Samples"))
    insts = num(d.get("Instructions Executed"))
    st = {k: num(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
    lines.append((cur, ln, r[1].strip(), samples, insts, st))
tot_s = sum(x[3] for x in lines)
tot_i = sum(x[4] for x in lines)
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
byfile = defaultdict(lambda: [0, 0])
for f, ln, src, s, i, st in lines:
    byfile[f][0] += s
    byfile[f][1] += i
for f, (s, i) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-16s samples %5.1f %%  instructions %5.1f %%" % (f, 100.0 * s / tot_s, 100.0 * i / tot_i))
print("top lines:")
for f, ln, src, s, i, st in sorted(lines, key=lambda x: -x[3])[:top_n]:
    worst = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("  %5.2f %%  inst %5.2f %%  %s:%d  %-70s %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, f, ln, src[:70],
                                                     " ".join("%s=%d" % (k[6:], v) for k, v in worst)))
if len(sys.argv) > 3:   # phase table: name:file:first-last,...
    print("phases:")
    for spec in sys.argv[3].split(","):
        name, f, rng = spec.split(":")
        a, b = (int(v) for v in rng.split("-"))
        s = sum(x[3] for x in lines if x[0] == f and a <= x[1] <= b)
        i = sum(x[4] for x in lines if x[0] == f and a <= x[1] <= b)
        agg = defaultdict(int)
        for x in lines:
            if x[0] == f and a <= x[1] <= b:
                for k, v in x[5].items():
                    agg[k] += v
        worst = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
        print("  %-14s samples %5.1f %%  instructions %5.1f %%   %s" % (name, 100.0 * s / tot_s, 100.0 * i / tot_i,
              " ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(1, s)) for k, v in worst)))
