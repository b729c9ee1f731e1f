#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report -> markdown table (for profiles/).

    python tools/ncu_summary.py REPORT.ncu-rep [KERNEL_REGEX]"""
import csv
import io
import re
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("smsp__inst_executed.sum", "warp-inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        short = re.sub(r"\(.*", "", name).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        cells = []
        for k, _ in COLS:
            if k in ix:
                v, u = r[ix[k]], units[ix[k]]
                try:
                    f = float(v)
                    v = f"{f:.3g}" if abs(f) < 1e6 else f"{f:.3e}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        print(f"| {short} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
