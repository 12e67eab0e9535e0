#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (nvdisasm --print-line-info output):
    cuobjdump -xelf all fused.o; nvdisasm --print-line-info fused.sm_100a.cubin > lines.txt
    python profiles/tools/sass_by_source.py lines.txt <mangled kernel substring> [name:file:first-last,...]"""
import re
import sys
from collections import defaultdict

txt = open(sys.argv[1]).read().split("\n")
key = sys.argv[2]
inside, cur = False, None
cnt = defaultdict(int)
for ln in txt:
    if ln.startswith("\t.text.") or ln.startswith(".text."):
        inside = key in ln
        continue
    if ln.lstrip().startswith(".section"):
        inside = key in ln and ".text." in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln) and cur:
        cnt[cur] += 1
tot = sum(cnt.values())
print("kernel", key, "static instructions with line info:", tot, "(%.0f KB)" % (tot * 16 / 1024))
byfile = defaultdict(int)
for (f, l), c in cnt.items():
    byfile[f] += c
for f, c in sorted(byfile.items(), key=lambda kv: -kv[1]):
    print("  %-24s %6d" % (f, c))
if len(sys.argv) > 3:
    for spec in sys.argv[3].split(","):
        name, f, rng = spec.split(":")
        a, b = (int(v) for v in rng.split("-"))
        c = sum(v for (ff, l), v in cnt.items() if ff == f and a <= l <= b)
        print("  %-16s %6d  %5.1f %%" % (name, c, 100.0 * c / tot))
else:
    for (f, l), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
        print("  %s:%d  %d" % (f, l, c))
