#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report -> markdown table (for profiles/).

    python tools/ncu_summary.py REPORT.ncu-rep [KERNEL_REGEX] [--json FRAMES SOURCE_NOTE]

--json: also rewrites profiles/alu.json (thread-instructions per frame, issue-slot utilisation, lanes per instruction
of the raster / backward kernels: bench.py's instruction roofline) and profiles/traffic.json (DRAM bytes per frame)
from this report, FRAMES = frames per launch of the captured run."""
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
    if "--json" in sys.argv:
        import json
        import os
        frames = int(sys.argv[sys.argv.index("--json") + 1])
        note = sys.argv[sys.argv.index("--json") + 2]
        names = [(r"k_raster<(\(bool\))?1>", "raster"), (r"k_backward<(\(bool\))?1, (\(bool\))?1>", "backward"),
                 (r"k_corr", "corr"), (r"k_setup_bin", "setup_bin"), (r"k_neg_maps", "neg_maps")]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        alu, traffic = {}, {"_note": note}
        for r in rows[2:]:
            key = next((v for k, v in names if re.search(k, r[ix["Kernel Name"]])), None)
            if key is None or key in alu:
                continue
            g = lambda m: float(r[ix[m]])  # noqa: E731
            gb = lambda m: float(r[ix[m]]) * scale.get(units[ix[m]], 1.0)  # noqa: E731
            warp_inst = g("smsp__inst_executed.sum")
            lanes = g("smsp__thread_inst_executed_per_inst_executed.ratio")
            alu[key] = {"thread_inst_per_frame": warp_inst * lanes / frames,
                        "warp_inst_per_frame": warp_inst / frames,
                        "issue_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "lanes": lanes, "source": note}
            traffic[key] = {"dram_bytes_per_frame": (gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")) / frames}
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        json.dump(alu, open(os.path.join(root, "profiles", "alu.json"), "w"), indent=1)
        json.dump(traffic, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
