"""Instruction-issue roofline of the cutout kernels from an .ncu-rep (`ncu -i rep --page raw --csv` piped in): warp instructions executed
per SM sub-partition and cycle vs the issue peak of 1, next to the DRAM throughput -- the measured form of "these kernels are ALU /
issue bound, not HBM bound".  Usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_issue_roof.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
def g(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return float("nan")
print("# %-34s %9s %12s %10s %10s %8s %8s %8s" % ("kernel", "time_us", "warp_inst", "inst/clk", "issue%", "dram%", "fma%", "xu%"))
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
    t = g(r, "gpu__time_duration.sum")
    tu = units[col["gpu__time_duration.sum"]]
    t_us = t / 1e3 if tu in ("ns", "nsecond") else (t if tu.startswith("us") else t * 1e3)
    inst = g(r, "smsp__inst_executed.sum")
    ipc = g(r, "smsp__inst_executed.avg.per_cycle_active") if "smsp__inst_executed.avg.per_cycle_active" in col else float("nan")
    print("%-36s %9.1f %12.0f %10.3f %10.1f %8.1f %8.1f %8.1f" % (name[:36], t_us, inst, ipc, g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
          g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
          g(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")))
