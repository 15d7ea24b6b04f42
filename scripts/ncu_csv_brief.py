"""Per-launch table from an `ncu --csv --metrics ...` log: time, DRAM bytes, L2 bytes, tensor-pipe share.
  python scripts/ncu_csv_brief.py gpurun_out/x.csv [name-filter]"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
hdr = rows[0]
ik, im, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = OrderedDict()
for r in rows[1:]:
    d.setdefault((r[iid], r[ik].replace('<unnamed>::', '').split('(')[0][:44]), {})[r[im]] = float(r[iv].replace(',', '') or 0)
for (i, k), v in d.items():
    if flt and flt not in k:
        continue
    t = v.get('gpu__time_duration.sum', 0) / 1e3
    dr = v.get('dram__bytes_read.sum', 0) + v.get('dram__bytes_write.sum', 0)
    print(f"{i:>4} {k:44s} {t:8.1f} us  dram {dr / 1e6:8.1f} MB  lts {v.get('lts__t_bytes.sum', 0) / 1e6:9.1f} MB  "
          f"tensor {v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):5.1f} %")
