"""Per-kernel SASS evidence of the Blackwell paths in libturtle_b200.so: counts of UTCHMMA / UTCQMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store), UTCBAR (tcgen05.commit), FHFMA, FFMA2.

  python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "turtlevsr_b200", "lib", "libturtle_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "FHFMA", "FFMA2", "HFMA2", "MUFU"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*$", "", name).replace("void ", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_all"] += 1
        for o in OPS:
            if op.startswith(o):
                cur[o] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a); instruction counts per kernel (static)")
print("kernel".ljust(46) + " ".join(o.rjust(8) for o in OPS) + "    total")
for name, c in counts.items():
    if not any(c[o] for o in OPS[:8]) and c["FHFMA"] == 0:
        continue
    print(name[:45].ljust(46) + " ".join(str(c[o]).rjust(8) for o in OPS) + f" {c['_all']:8d}")
    for o in OPS:
        total[o] += c[o]
print("TOTAL".ljust(46) + " ".join(str(total[o]).rjust(8) for o in OPS))
print(f"# {len(counts)} kernels in the library; rows without tcgen05 / TMA / FHFMA instructions omitted")
