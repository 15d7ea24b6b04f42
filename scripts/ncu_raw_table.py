"""Per-launch table out of `ncu -i report.ncu-rep --page raw --csv` (time, DRAM bytes, pipe / issue / occupancy figures).

  python scripts/ncu_raw_table.py raw.csv > profiles/<tag>.txt
"""
import csv
import sys

WANT = [("us", "gpu__time_duration.sum"), ("dramR MB", "dram__bytes_read.sum"), ("dramW MB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1tex%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("smemKB", "launch__shared_mem_per_block_dynamic")]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}
TSCALE = {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none, raw page of {path.split('/')[-1]}: one row per profiled launch")
    print("# tensor% = sm__pipe_tensor_cycles_active, issue% = smsp__issue_active, warps% = sm__warps_active (all % of peak)")
    print("kernel".ljust(44), " ".join(w[0].rjust(9) for w in WANT))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:43]
        vals = []
        for short, m in WANT:
            if m not in hdr:
                vals.append("-".rjust(9))
                continue
            i = hdr.index(m)
            try:
                f = float(r[i].replace(",", ""))
                if "MB" in short:
                    f *= SCALE.get(units[i], 1)
                elif short == "us":
                    f *= TSCALE.get(units[i], 1)
                elif short == "smemKB":
                    f *= {"byte": 1 / 1024, "Kbyte": 1, "Mbyte": 1024}.get(units[i], 1)
                vals.append(f"{f:9.1f}")
            except ValueError:
                vals.append(r[i][:9].rjust(9))
        print(name.ljust(44), " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
