#!/usr/bin/env python
"""Summarise an .ncu-rep (full set) into a small text table for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.txt"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "warp_inst")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, units = rows[0], rows[1]
    ki = H.index("Kernel Name")
    cols = [(H.index(m), n) for m, n in WANT if m in H]
    print(f"# ncu --set full --clock-control none summary of {rep}")
    print("# units: " + ", ".join(f"{n}[{units[i]}]" for i, n in cols))
    print("%-44s" % "kernel" + "".join("%12s" % n for _, n in cols))
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        print("%-44s" % name[:44] + "".join("%12s" % (r[i][:11]) for i, _ in cols))


if __name__ == "__main__":
    main()
