"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py <tag> launches.csv [name=report.ncu-rep ...]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
]


def launches(tag, path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({os.path.basename(path)})\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us total\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:48s} launches={v[0]:4d}  {v[1] / 1e3:10.1f} us  {100 * v[1] / tot:5.1f}%\n")


def report(tag, name, path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, f"{tag}_{name}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on   ({os.path.basename(path)})\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"  {m:70s} {r[i]:>16s} {units[i]}\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    for kv in sys.argv[3:]:
        n, p = kv.split("=")
        report(tag, n, p)
