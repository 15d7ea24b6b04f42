"""Compact per-launch table of an .ncu-rep (raw page): time, DRAM bytes, issue/stall summary."""
import csv, subprocess, sys
path = sys.argv[1]
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
def col(n):
    return hdr.index(n) if n in hdr else None
cols = [("us", "gpu__time_duration.sum"), ("rdMB", "dram__bytes_read.sum"), ("wrMB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("inst", "smsp__inst_executed.sum")]
units = rows[1]
print("kernel".ljust(44), "grid".ljust(12), " ".join(c[0].rjust(9) for c in cols))
for r in rows[2:]:
    name = r[col("Kernel Name")].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:43]
    vals = []
    for short, m in cols:
        i = col(m)
        v = r[i] if i is not None else ""
        u = units[i] if i is not None else ""
        try:
            f = float(v.replace(",", ""))
            if short in ("rdMB", "wrMB"):
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            if short == "us":
                f *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
            vals.append(f"{f:9.1f}" if f < 1e7 else f"{f:9.2e}")
        except ValueError:
            vals.append(v[:9].rjust(9))
    print(name.ljust(44), r[col("Grid Size")].replace(" ", "").ljust(12), " ".join(vals))
