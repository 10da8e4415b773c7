"""Launch list -> per-kernel table.  Input: the CSV log of
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X.csv <command>
Usage: python tools/ncu_launch_table.py X.csv "<command line>" > profiles/<name>.txt ; also prints the GEMM's average DRAM bytes per launch as JSON on stderr."""
import csv, json, sys
from collections import OrderedDict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path) if l.startswith('"')]
rd = csv.DictReader(lines)
launch = OrderedDict()  # ID -> dict
for r in rd:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"], "t": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["t"] = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
    else:
        b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d["rd" if "read" in r["Metric Name"] else "wr"] = b
agg = OrderedDict()
for d in launch.values():
    n = d["name"].split("(")[0]
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d["t"]; a[2] += d["rd"]; a[3] += d["wr"]
total = sum(a[1] for a in agg.values())
print("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:<our kernels>")
print("# %s" % cmd)
print("# per-launch times are cold-cache and serialised: compare SHARES.  our kernels total %.1f us over %d launches" % (total, sum(a[0] for a in agg.values())))
print("# %6s %5s %10s %9s %12s %12s  kernel" % ("share", "calls", "total_us", "avg_us", "dram_rd_MB/l", "dram_wr_MB/l"))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%7.2f%% %5d %10.1f %9.1f %12.2f %12.2f  %s" % (100 * a[1] / total, a[0], a[1], a[1] / a[0], a[2] / a[0] / 1e6, a[3] / a[0] / 1e6, n))
g = [(a[0], a[2] + a[3]) for n, a in agg.items() if "gemm_bf16_tn" in n]
if g:
    sys.stderr.write(json.dumps({"launches": sum(x[0] for x in g), "avg_dram_bytes_per_launch": sum(x[1] for x in g) / sum(x[0] for x in g)}) + "\n")
