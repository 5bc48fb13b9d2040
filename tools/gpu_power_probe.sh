# Power draw / SM clock under the C3 step (is the step power-capped, and where?).  Output: gpurun_out/r02_power_probe.txt
mkdir -p gpurun_out
nvidia-smi -q -d POWER | grep -i "power limit\|power draw\|default" | head -8 > gpurun_out/r02_power_probe.txt
nvidia-smi --query-gpu=power.draw,power.draw.instant,clocks.sm,clocks_throttle_reasons.sw_power_cap,temperature.gpu --format=csv,noheader -lms 100 > gpurun_out/power_samples.csv &
SMI=$!
sleep 2
python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-ctc --no-e2e > gpurun_out/power_bench.json 2>/dev/null
sleep 1
kill $SMI
python - <<'PY' >> gpurun_out/r02_power_probe.txt
import json
rows = [l.strip().split(", ") for l in open("gpurun_out/power_samples.csv") if l.strip()]
d = json.load(open("gpurun_out/power_bench.json"))
print("bench: %.2f ms per step, %d steps" % (d["ms_per_step"], d["steps"]))
busy = [r for r in rows if float(r[0].split()[0]) > 400]
idle = [r for r in rows if float(r[0].split()[0]) <= 400]
def col(rs, i): return sorted(float(r[i].split()[0]) for r in rs)
for name, rs in (("under load", busy), ("idle", idle)):
    if rs:
        p, pi, c = col(rs, 0), col(rs, 1), col(rs, 2)
        print("%s: %d samples; power.draw W min/median/max %.0f/%.0f/%.0f; instant %.0f/%.0f/%.0f; SM MHz min/median/max %.0f/%.0f/%.0f; sw_power_cap active in %d; temp %s C"
              % (name, len(rs), p[0], p[len(p) // 2], p[-1], pi[0], pi[len(pi) // 2], pi[-1], c[0], c[len(c) // 2], c[-1],
                 sum(1 for r in rs if r[3].strip().lower().startswith("active")), rs[len(rs) // 2][4]))
PY
cat gpurun_out/r02_power_probe.txt
