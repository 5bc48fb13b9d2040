"""Per-GEMM tensor throughput from `ncu --metrics ... -k regex:gemm_bf16_tcgen05 --csv` (tools/gpu_final_round2.sh):
tcgen05.mma instruction counts -> flops -> TFLOP/s per launch class, against the burst peak of the SMs the launch used.
Usage: python tools/gemm_util_summary.py gpurun_out/gemm_launch_metrics.csv > profiles/r01_gemm_tensor_utilisation.txt"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(_pk)) if os.path.exists(_pk) else {}          # driver-written per pod; fallback below


def find(d, *keys):
    for k, v in d.items():
        if isinstance(v, dict):
            r = find(v, *keys)
            if r is not None:
                return r
        elif any(x in k.lower() for x in keys) and isinstance(v, (int, float)):
            return v
    return None


burst = peak.get("bf16_tflops") or find(peak, "burst") or 1633.2
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ID, NAME, GRID, MN, UNIT, VAL = (h.index(x) for x in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value"))
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= VAL:
        continue
    d = L.setdefault(r[ID], {"name": r[NAME], "grid": int(re.findall(r"\d+", r[GRID])[0])})
    try:
        d[r[MN]] = (float(r[VAL].replace(",", "")), r[UNIT])
    except ValueError:
        pass
WHAT = {("256, 0, 0, 0", 245760): "projection tail X*Wx (64 % of the frames, capped grid beside the forward recurrence)",
        ("256, 0, 0, 0", 138240): "projection head X*Wx (36 % of the frames, whole chip)",
        ("256, 0, 1, 1", 138240): "dX early rows, first cut (capped grid beside BPTT)", ("256, 0, 1, 1", 262144): "dX early rows",
        ("256, 0, 1, 1", 115712): "dX late rows (whole chip, on the serial chain)",
("256, 1, 1, 0", 768000): "wgrad dWx (K = 96000, capped grid)", ("256, 0, 1, 1", 768000): "dgrad dX = dG*Wx (bf16 out, fused dropout)",
        ("256, 0, 1, 0", 48000): "dM = dH*W_proj^T (fp32 out)", ("256, 0, 0, 2", 48000): "h = m*W_proj (fp16 out, fused dropout)",
        ("128, 1, 1, 0", 383744): "wgrad dW' (split-K, capped grid)", ("128, 1, 1, 0", 96000): "wgrad dW_proj (split-K, capped grid)"}
cls = collections.OrderedDict()
for d in L.values():
    t = re.search(r"<([^>]*)>", d["name"]).group(1)
    dur = d["gpu__time_duration.sum"]
    us = dur[0] * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(dur[1], 1e-3)
    inst = int(d["sm__inst_executed_pipe_tensor.sum"][0])
    # (sm__pipe_tensor_cycles_active_realtime is only reported inside ncu's sections, not as a --metrics name on this ncu: the
    # column shows the SM-throughput percentage instead; the tensor-pipe figure of the big launches is in r01_ncu_full_summary.txt)
    tp = d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", (float("nan"),))[0]
    c = cls.setdefault((t, d["grid"], inst), [0, 0.0, 0.0])
    c[0] += 1
    c[1] += us
    c[2] += tp
print("tcgen05.mma instruction counts (ncu sm__inst_executed_pipe_tensor.sum) -> achieved tensor throughput per GEMM launch class of one C3 training step")
print("flops = instructions x 2*128*BN*16; per-SM peak = measured cuBLAS bf16 burst %.1f TFLOP/s / 148 SMs (MEASURED_PEAKS.json); grid = persistent CTAs = SMs used" % burst)
print("(weight-gradient GEMMs run on a capped grid of 80 beside the BPTT clusters; projection tails on 84 beside the forward recurrence);")
print("sm% = ncu sm__throughput.avg.pct_of_peak_sustained_elapsed over ALL 148 SMs; times are ncu's (cold cache, serialised)")
print("%-18s %5s %9s %4s %9s %9s %15s %8s  %s" % ("template<BN,A,B,C>", "grid", "mma inst", "n", "avg us", "TFLOP/s", "% of SMs' peak", "sm%", "what"))
for (t, grid, inst), (n, us, tp) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    if us / n < 20:
        continue
    bn = int(t.split(",")[0])
    tf = inst * 2.0 * 128 * bn * 16 / (us / n * 1e-6) / 1e12
    print("<%-16s> %5d %9d %4d %9.1f %9.1f %14.1f%% %8.1f  %s" % (t, grid, inst, n, us / n, tf, 100 * tf / (burst * min(grid, 148) / 148), tp / n,
                                                                  WHAT.get((t, inst), "")))
