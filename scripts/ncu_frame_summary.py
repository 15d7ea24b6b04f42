"""One steady-state frame out of an ncu launch list (gpu__time_duration + dram bytes per launch).

  python scripts/ncu_frame_summary.py <tag> gpurun_out/launches.csv [frame-index, default 3]

The list is cut between two consecutive pack_frame_kernel launches (= one forward of the arch).  Writes
profiles/<tag>_launches_summary.txt (per-kernel launches, serialised cold-cache time, share, DRAM bytes) and
profiles/<tag>_traffic.json (per-kernel DRAM read+write bytes per frame and per launch; bench.py's roofline.traffic).
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
# C-ABI entry point (bench.py's per-kernel names) -> device kernels it launches
ENTRY = {
    "turtle_gemm": ("gemm_tc2_kernel", "gemm_tc_kernel", "gemm_f32_kernel"),
    "turtle_dwconv3x3": ("dwconv16_kernel", "dwconv_tma_kernel", "dwconv3x3_kernel"),
    "turtle_layernorm": ("layernorm_vec_kernel", "layernorm_kernel"),
    "turtle_sab_aggregate": ("sab_aggregate",),
    "turtle_sab_aggregate_tc": ("sab_agg_tc_kernel", "sab_wd_build_kernel", "sab_far_add_kernel"),
    "turtle_chan_gram": ("gram_tc_kernel", "gram64_kernel", "chan_gram_kernel"),
}


def main(tag, path, which=3):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mn, mv, mu, idc = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    L = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        d = L.setdefault(r[idc], {"name": r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")})
        v = float(r[mv].replace(",", ""))
        if "time" in r[mn]:
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[mu], 1)
        else:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[mu], 1)
        d[r[mn]] = v
    ls = list(L.values())
    pf = [i for i, d in enumerate(ls) if d["name"].startswith("pack_frame")]
    # frame `which` of the capture (default 3 = bench.py's first timed frame after --warmup 3: weights packed,
    # history rings full); earlier frames pack weights / run with a part-filled history
    which = min(which, len(pf) - 2)
    fr = ls[pf[which]:pf[which + 1]]
    agg = collections.OrderedDict()
    for d in fr:
        x = agg.setdefault(d["name"], [0, 0.0, 0.0])
        x[0] += 1
        x[1] += d["gpu__time_duration.sum"]
        x[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(v[1] for v in agg.values())
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f"{tag}_launches_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none"
                f"  ({os.path.basename(path)})\n")
        f.write("# one steady-state 1280x720 frame (between two pack_frame launches); per-launch times are cold-cache and\n"
                "# serialised: compare SHARES with bench.py's per_kernel_ms, not absolutes\n")
        f.write(f"# {len(fr)} launches, {tot / 1e3:.2f} ms total, {sum(v[2] for v in agg.values()) / 1e9:.1f} GB DRAM traffic\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:50s} launches={v[0]:4d} {v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%  dram {v[2] / 1e9:7.2f} GB"
                    f"  {v[2] / v[1] / 1e3:6.0f} GB/s\n")
    traffic = {}
    for entry, kernels in ENTRY.items():
        n = sum(v[0] for k, v in agg.items() if k.startswith(kernels))
        by = sum(v[2] for k, v in agg.items() if k.startswith(kernels))
        us = sum(v[1] for k, v in agg.items() if k.startswith(kernels))
        if n:
            traffic[entry] = {"launches_per_frame": n, "dram_bytes_per_frame": by, "dram_bytes_per_launch": by / n,
                              "us_per_frame_serialised": us}
    json.dump({"source": os.path.basename(path), "frame_launches": len(fr), "kernels": traffic},
              open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
    print(open(os.path.join(OUT, f"{tag}_launches_summary.txt")).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 3)
