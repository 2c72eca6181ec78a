#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into a small per-launch table: python scripts/ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tc_pipe_%"),
    ("sm__inst_executed_pipe_tc.sum", "tc_inst"),
    ("smsp__inst_executed.sum", "inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dsmem_B"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(head)}
    cols = [(m, s) for m, s in METRICS if m in idx]
    lines = ["| # | kernel | " + " | ".join(f"{s} [{units[idx[m]]}]" if units[idx[m]] else s for m, s in cols) + " |",
             "|---|---|" + "---|" * len(cols)]
    for k, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]]
        name = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        lines.append(f"| {k} | `{name}` | " + " | ".join(r[idx[m]] for m, _ in cols) + " |")
    out = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(out)
    else:
        print(out)


if __name__ == "__main__":
    main()
