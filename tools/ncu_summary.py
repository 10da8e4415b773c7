"""Per-launch table from an .ncu-rep (`ncu -i rep --page raw --csv` piped in): time, DRAM bytes, achieved GB/s, dram%/sm% of ncu's
peaks, registers, achieved occupancy.  Usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
def g(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default
units = rows[1]
print("# %-44s %9s %10s %10s %8s %6s %6s %5s %6s" % ("kernel", "time_us", "dram_rd_MB", "dram_wr_MB", "GB/s", "dram%", "sm%", "regs", "occ%"))
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
    t = g(r, "gpu__time_duration.sum")
    tu = units[col["gpu__time_duration.sum"]]
    t_us = t / 1e3 if tu in ("ns", "nsecond") else (t if tu.startswith("us") else t * 1e3 if tu.startswith("ms") else t)
    def mb(name_):
        v, u = g(r, name_), units[col[name_]]
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    rd, wr = mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum")
    print("%-46s %9.1f %10.2f %10.2f %8.0f %6.1f %6.1f %5d %6.1f" % (name[:46], t_us, rd, wr, (rd + wr) / t_us * 1e3 if t_us else 0,
          g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
          int(g(r, "launch__registers_per_thread")), g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
