"""Per-kernel census of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised times: the SHARE per
kernel is what compares with the live CUDA-event breakdown in the bench line).
Usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/r01_ncu_launch_summary_c3.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
NAME, UNIT, VAL, MN = h.index("Kernel Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("Metric Name")
acc = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= VAL or r[MN] != "gpu__time_duration.sum":
        continue
    us = float(r[VAL].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[UNIT], 1e-3)
    name = re.sub(r"\(.*", "", r[NAME]).strip()
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in acc.values())
print("kernel,launches,total_us,share")
for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.4f" % (k.replace(",", ";"), n, us, us / tot))
