#!/usr/bin/env python
"""Summarise an .ncu-rep (full set) into a small text table for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--traffic-json profiles/ncu_traffic.json] > profiles/rNN_xxx.txt
--traffic-json also writes DRAM bytes (read + written) per launch of every kernel, stamped with the md5 of
csrc/cmfd_kernels.cu: bench.py reports it as roofline.traffic only while that source is unchanged."""
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
    traffic = {}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        print("%-44s" % name[:44] + "".join("%12s" % (r[i][:11]) for i, _ in cols))
        try:
            ir, iw = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")
            b = float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0)
            short = name.split("<")[0]
            traffic.setdefault(short, []).append(b)
        except (ValueError, IndexError):
            pass
    if "--traffic-json" in sys.argv:
        import hashlib, json, os
        root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
        md5 = hashlib.md5(open(os.path.join(root, "adpres_b200", "csrc", "cmfd_kernels.cu"), "rb").read()).hexdigest()
        out_path = sys.argv[sys.argv.index("--traffic-json") + 1]
        json.dump({"cmfd_kernels_md5": md5, "source": os.path.basename(rep),
                   "kernels": {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v)} for k, v in traffic.items()}},
                  open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
