#!/usr/bin/env python
"""SASS opcode histograms of every kernel in libdynhor_b200.so (cuobjdump -sass) -> profiles/: tracked evidence of
what the built code is made of (tcgen05 = UTCHMMA / LDTM / UTCBAR, TMA = UTMALDG / UBLKCP, mbarrier = SYNCS, ...).

    python tools/sass_hist.py [LIB.so] > profiles/rN_sass_histograms.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dynhor_b200/libdynhor_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMAPF", "HMMA", "REDUX",
       "ATOMS", "ATOMG", "RED", "MUFU", "F2I", "FFMA2", "BAR", "WARPSYNC", "SHFL", "VOTE", "LDG", "STG", "LDS", "STS")
print(f"# cuobjdump -sass {lib}: static instruction counts per kernel (sm_100a)\n")
for k, h in hist.items():
    total = sum(h.values())
    print(f"== {k}   ({total} instructions)")
    keys = {}
    for op, n in h.items():
        base = op.split(".")[0]
        if base in KEY:
            keys[base] = keys.get(base, 0) + n
    print("   key mnemonics: " + ", ".join(f"{a} {b}" for a, b in sorted(keys.items(), key=lambda kv: -kv[1])))
    top = sorted(h.items(), key=lambda kv: -kv[1])[:24]
    print("   top opcodes:   " + ", ".join(f"{a} {b}" for a, b in top))
    print()
