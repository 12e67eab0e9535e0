#!/usr/bin/env python
"""SASS opcode histogram per kernel of libkmpc.so's translation units (cuobjdump -sass on csrc/_obj/*.o).
    python profiles/tools/sass_histogram.py > profiles/r2/sass_opcodes.txt
Proves which hardware paths a kernel uses: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG
(cp.async.bulk.tensor), UBLKCP (cp.async.bulk), DMMA (mma.sync f64), SYNCS (mbarrier)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
OBJ = os.path.join(ROOT, "koopman_online_updated_mpc_b200", "csrc", "_obj")
KEY = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "DMMA", "DFMA", "HMMA", "F2FP",
       "ELECT", "BAR", "LDS", "STS", "LDG", "STG", "SHFL", "MUFU")
for name in sorted(os.listdir(OBJ)):
    if not name.endswith(".o"):
        continue
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, name)], capture_output=True, text=True).stdout
    kern, hist = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            hist[kern][m.group(1)] += 1
    for kern, h in hist.items():
        n = sum(h.values())
        if n < 200:
            continue
        keys = " ".join("%s=%d" % (k, h[k]) for k in KEY if h.get(k))
        top = " ".join("%s=%d" % kv for kv in h.most_common(8))
        print("%s :: %s\n    %d instructions (%.0f KB)\n    key: %s\n    top: %s" % (name, kern, n, n * 16 / 1024, keys, top))
