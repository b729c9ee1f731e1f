#!/usr/bin/env python
"""Per-source-line summary of an ncu report (--set full --import-source on, built with -lineinfo).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [LIB.so] [--top N]

Joins `ncu --page source --print-source sass` (per-SASS-instruction counters) with `nvdisasm -gi` line info of the
kernel's cubin by instruction order, then aggregates warp-instructions, thread-instructions and stall samples per
source line (innermost inlined location)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_rows(report, kernel):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass",
                          "--kernel-name", f"regex:{kernel}"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and row and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] and row and row[0].startswith("0x"):
            cur["rows"].append(row)
    return blocks


def line_table(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    tables = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        fn, chain, fresh, rows = None, [], True, []
        for ln in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                if fn and rows:
                    tables[fn] = rows
                fn, rows, chain, fresh = m.group(1), [], [], True
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                if fresh:          # first line-info comment after an instruction starts a new inline chain
                    chain, fresh = [], False
                chain.append((os.path.basename(m.group(1)), int(m.group(2))))   # innermost first
                continue
            if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                fresh = True
                if not chain:
                    rows.append(None)
                else:
                    outer_first = chain[::-1]
                    rows.append(outer_first[min(LEVEL, len(outer_first) - 1)])
        if fn and rows:
            tables[fn] = rows
    return {k: v for k, v in tables.items() if re.search(kernel, k)}


LEVEL = 1  # inline depth used for attribution: 0 = line in the kernel body, 1 = first inlined callee, ...


def main():
    global LEVEL
    if "--level" in sys.argv:
        LEVEL = int(sys.argv[sys.argv.index("--level") + 1])
    report, kernel = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "dynhor_b200/libdynhor_b200.so"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    tabs = line_table(lib, kernel)
    for blk in sass_rows(report, kernel):
        hdr = blk["hdr"]
        ix = {h: i for i, h in enumerate(hdr)}
        # match the cubin function with the same instruction count
        cand = [v for v in tabs.values() if len(v) == len(blk["rows"])]
        print(f"== {blk['name']}  ({len(blk['rows'])} SASS instructions)")
        if not cand:
            print("   no cubin function with matching instruction count:", {k: len(v) for k, v in tabs.items()})
            continue
        lines = cand[0]
        agg = defaultdict(lambda: [0, 0, 0])
        tot = [0, 0, 0]
        for row, loc in zip(blk["rows"], lines):
            vals = [int(float(row[ix["Instructions Executed"]] or 0)),
                    int(float(row[ix["Thread Instructions Executed"]] or 0)),
                    int(float(row[ix["# Samples"]] or 0))]
            for i in range(3):
                agg[loc][i] += vals[i]
                tot[i] += vals[i]
        print(f"   total warp-inst {tot[0]:.3e}  thread-inst {tot[1]:.3e}  avg lanes {tot[1] / max(tot[0], 1):.1f}  "
              f"samples {tot[2]}")
        if "--regions" in sys.argv:
            # --regions FILE:a-b=name,FILE:c-d=name,...  -> one row per named line range (+ the rest per file)
            spec = sys.argv[sys.argv.index("--regions") + 1]
            regs = []
            for part in spec.split(","):
                rng, name = part.split("=")
                f, ab = rng.split(":")
                a, b = ab.split("-")
                regs.append((f, int(a), int(b), name))
            ragg = defaultdict(lambda: [0, 0, 0])
            for loc, v in agg.items():
                key = "other" if loc is None else f"rest of {loc[0]}"
                if loc is not None:
                    for f, a, b, name in regs:
                        if loc[0] == f and a <= loc[1] <= b:
                            key = f"{name} {a}-{b}"
                            break
                for i in range(3):
                    ragg[key][i] += v[i]
            print("   region                                   warp-inst%  lanes  samples%")
            for key, v in sorted(ragg.items(), key=lambda kv: -kv[1][0]):
                print(f"   {key:40s} {100.0 * v[0] / max(tot[0], 1):8.1f}  {v[1] / max(v[0], 1):5.1f}  "
                      f"{100.0 * v[2] / max(tot[2], 1):8.1f}")
            continue
        if "--dump" in sys.argv:
            for loc, v in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
                name = f"{loc[0]}:{loc[1]}" if loc else "?"
                print(f"   {name:24s} {v[0]:12d} {v[1] / max(v[0], 1):5.1f} {v[2]:8d}")
            continue
        print("   file:line                warp-inst%  lanes  samples%")
        for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
            name = f"{loc[0]}:{loc[1]}" if loc else "?"
            print(f"   {name:24s} {100.0 * v[0] / max(tot[0], 1):8.1f}  {v[1] / max(v[0], 1):5.1f}  "
                  f"{100.0 * v[2] / max(tot[2], 1):8.1f}")


if __name__ == "__main__":
    main()
