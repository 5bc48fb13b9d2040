"""DRAM bytes per launch of the dominant kernel families from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` pass over the bench command.
Usage: python tools/traffic_summary.py gpurun_out/r02_traffic.csv profiles/r02_ncu_traffic_raw.json profiles/r02_ncu_traffic.json"""
import collections
import csv
import json
import re
import sys

sys.path.insert(0, ".")
import bench

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ID, NAME, UNIT, VAL, MN = h.index("ID"), h.index("Kernel Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("Metric Name")
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= VAL:
        continue
    d = per.setdefault(r[ID], {"name": re.sub(r"\(.*", "", r[NAME]).replace("lcb::", "").strip()})
    v = float(r[VAL].replace(",", ""))
    if r[MN].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[UNIT], 1)
        d["bytes"] = d.get("bytes", 0.0) + v
    elif r[MN] == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[UNIT], 1e-3)
raw = collections.OrderedDict()
for d in per.values():
    a = raw.setdefault(d["name"], {"launches": 0, "bytes": 0.0, "us": 0.0})
    a["launches"] += 1; a["bytes"] += d.get("bytes", 0.0); a["us"] += d.get("us", 0.0)
out_raw = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"], "us_per_launch_under_ncu": v["us"] / v["launches"]}
           for k, v in raw.items()}
json.dump(out_raw, open(sys.argv[2], "w"), indent=1)


def fam(pred):
    ks = [k for k in raw if pred(k)]
    n = sum(raw[k]["launches"] for k in ks)
    return (sum(raw[k]["bytes"] for k in ks) / n, n) if n else (None, 0)


src = ("%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over the kernels of ~1.3 C3 training steps of "
       "`python bench.py --steps 1 --warmup 3`, tools/gpu_job.sh; mean per launch)" % sys.argv[2])
f, nf = fam(lambda k: "lstm_rec_fwd" in k)
b, nb = fam(lambda k: "lstm_rec_bwd" in k)
g, ng = fam(lambda k: "gemm_bf16_tcgen05" in k)
cs, _ = fam(lambda k: "ctc_softmax" in k)
cl, _ = fam(lambda k: "ctc_lattice" in k)
w = bench.WORKLOADS["c3"]
fam_out = {"lstm_rec_fwd": {"dram_bytes_per_launch": f, "launches_seen": nf, "source": src + "; one flow-controlled launch per layer"},
           "lstm_rec_bwd": {"dram_bytes_per_launch": b, "launches_seen": nb, "source": src + "; top layer: three launches, others one"},
           "ctc_loss_grad": {"dram_bytes_per_launch": (cs or 0) + (cl or 0), "source": src + "; ctc_softmax_kernel + ctc_lattice_kernel of one call"},
           "gemm": {"dram_bytes_per_launch": g, "launches": ng, "source": src + "; all gemm_bf16_tcgen05_kernel launches"}}
json.dump({w["desc"]: fam_out}, open(sys.argv[3], "w"), indent=1)
print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in fam_out.items()}))
