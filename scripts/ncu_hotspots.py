"""Hot spots of one kernel from an .ncu-rep's source page: samples / executed instructions per CUDA source line, the
dominant stall reasons, shared-memory bank conflict excess.   python scripts/ncu_hotspots.py report.ncu-rep [kernel-id] [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kid is not None:
    cmd += ["--kernel-id", f":::{kid}"] if False else []
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the dump is a sequence of per-kernel blocks; each block: "File Path", "Function Name", header row, lines...
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Function Name":
        cur = {"fn": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r and r[0] not in ("File Path", "Kernel Name"):
        cur["rows"].append(r)
sel = int(kid) if kid is not None else 0
names = []
for b in blocks:
    if b["fn"] not in names:
        names.append(b["fn"])
fn = names[sel]
print("kernel:", fn)
agg = collections.OrderedDict()
tot_s = tot_i = 0
for b in blocks:
    if b["fn"] != fn:
        continue
    h = b["hdr"]
    si, ii = h.index("# Samples"), h.index("Instructions Executed")
    ex, wf = h.index("L1 Wavefronts Shared Excessive"), h.index("L1 Wavefronts Shared")
    st = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    for r in b["rows"]:
        if not r[0]:           # SASS row
            continue
        key = (b.get("file", ""), r[0], r[1].strip()[:90])
        d = agg.setdefault(key, dict(s=0, i=0, ex=0, wf=0, stalls=collections.Counter()))
        try:
            d["s"] += int(r[si]); d["i"] += int(r[ii]); d["ex"] += int(r[ex]); d["wf"] += int(r[wf])
        except ValueError:
            continue
        for i in st:
            try:
                d["stalls"][h[i]] += int(r[i])
            except ValueError:
                pass
tot_s = sum(d["s"] for d in agg.values())
tot_i = sum(d["i"] for d in agg.values())
print(f"total samples {tot_s}, warp instructions executed {tot_i}")
allst = collections.Counter()
for d in agg.values():
    allst.update(d["stalls"])
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(1, sum(allst.values())):.1f}%" for k, v in allst.most_common(8)))
for (f, ln, src), d in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top]:
    stl = ", ".join(f"{k[6:]} {v}" for k, v in d["stalls"].most_common(3))
    print(f"{100 * d['s'] / tot_s:5.1f}% smp {100 * d['i'] / tot_i:5.1f}% inst  line {ln:>4}  bank-excess {d['ex']:>9}/{d['wf']:<9} {src[:70]:70s} | {stl}")
