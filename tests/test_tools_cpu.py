"""The summarising tools run on the committed ncu exports (so the files under profiles/ can be regenerated from the raw CSVs) and
the committed bench lines carry every key the measurement contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def _run(tool, *args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool)] + list(args), capture_output=True, text=True, check=True).stdout


def test_gemm_utilisation_summary_regenerates():
    out = _run("gemm_util_summary.py", os.path.join(P, "r01_gemm_launch_metrics.csv"))
    ref = open(os.path.join(P, "r01_gemm_tensor_utilisation.txt")).read()
    # same launch classes in the same order (the percentages depend on MEASURED_PEAKS.json, which the driver rewrites per pod)
    cls = lambda txt: [ln.split(">")[0] + ln.split(">")[1].split()[1] for ln in txt.splitlines() if ln.startswith("<")]
    assert cls(out) == cls(ref) and len(cls(out)) > 10
    assert "projection tail" in out and "nan" not in out


def test_bench_lines_carry_the_contract_keys():
    for name in ("r01_bench_c3_1gpu.json", "r01_bench_c3_2gpu.json", "r01_bench_c3_4gpu.json", "r01_bench_c3_8gpu.json"):
        d = json.load(open(os.path.join(P, name)))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (name, k)
        assert d["metric"] == "train_frames_per_sec" and d["scaling"] == "weak" and d["gpu_launches"] > 0
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert abs(d["value"] - d["config"]["frames_per_step_global"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    one = json.load(open(os.path.join(P, "r01_bench_c3_1gpu.json")))
    assert one["cpu_baseline"]["kind"] == "port" and one["cpu_baseline"]["cores"] >= 1 and one["roofline"]["traffic"]
    ref = json.load(open(os.path.join(P, "r01_bench_c3_reference_arm.json")))
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["metric"] == one["metric"]
    assert ref["config"]["workload"] == one["config"]["workload"]
