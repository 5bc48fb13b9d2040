"""Condense `ncu --page raw --csv` exports into the handful of metrics DESIGN.md / profiles/ quote.
Usage: python tools/ncu_summary.py name=path.csv [name=path.csv ...]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput % of peak"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "cycles per issued instruction (per warp)"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_dim_x", "cluster size"),
    ("launch__registers_per_thread", "registers / thread"),
    ("lts__t_bytes.sum", "L2 bytes"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            print("== %s: no launches captured" % name)
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, zip(r, units)))
            kn = d.get("Kernel Name", ("?", ""))[0]
            print("== %s: %s" % (name, kn[:110]))
            for k, label in KEYS:
                for h in hdr:
                    if h == k or h.endswith("." + k):
                        v, u = d[h]
                        print("   %-44s %s %s" % (label, v, u))
                        break
            stalls = []
            for h in hdr:
                if STALL in h and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(d[h][0]), h.split(STALL)[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("   top stalls (warps per issue)                 " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:5]))


if __name__ == "__main__":
    main()
