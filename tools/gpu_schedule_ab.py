"""Interleaved A/B of the schedule options on the C3 training step (one box, medians over rounds: the power-capped clock drifts by
more than some of the effects).  Historical first purpose: step time against the share of the idle SMs the side-stream GEMMs may use
(what a forward recurrence on 96 instead of 64 SMs would leave to the projection GEMMs beside it)."""
import sys
import json
import torch
sys.path.insert(0, ".")
import bench
from lstm_ctc_b200.model import AcousticModel

w = bench.WORKLOADS["c3"]
dev = torch.device("cuda:0")
model = AcousticModel(bench.nnet_config(w, 0.9), dev, seed=1234)
x_h, lens_h, y_h = bench.synth_batch(w, 777)
x, lens, y = x_h.to(dev), lens_h.to(dev), y_h.to(dev)


def step():
    model.loss_and_grad(x, lens, y, check_labels=False, seq_len_host=HOST[0])
    model.optimizer_step("adam", 4e-4, clip_norm=5.0, l2_decay_weight=1e-5)


def timed(n=12, warm=2):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


HOST = [lens_h]
from lstm_ctc_b200 import _lib
L = _lib.lib()
import statistics
HOST[0] = lens_h


def apply(cfg):
    L.lcb_debug_fwd_layout(1 if cfg.get("layout") == "paired" else 0)
    HOST[0] = lens_h if cfg.get("host", True) else None
    model.enc.fwd_flow_control = cfg.get("flow", True)
    model.enc.bwd_progress = cfg.get("pg", True)
    model.enc.bwd_early_fracs = cfg.get("early", [0.67, 0.85])
    model.enc.flow_fracs = cfg.get("fracs", [0.3, 0.55, 0.8])
    model.top_overlap = cfg.get("top", True)
    model.enc.fwd_hproj_fracs = cfg.get("hfracs", [0.6, 0.85])
    model.enc.g_half = cfg.get("g_half", True)
    model.enc.bf16_twins = cfg.get("twins", True)
    # timing-only experiment: no conversion pass at all (stale bf16 copies -> WRONG gradients; the time is what is measured)
    import lstm_ctc_b200.blstm as _b
    if not hasattr(_b, "_to_bf16_real"):
        _b._to_bf16_real = _b._to_bf16
    _b._to_bf16 = (lambda src, dst: dst) if cfg.get("skip_bf16_conv", False) else _b._to_bf16_real


# (earlier visits compared the chunk fractions: profiles/r02_flow_fracs_ab.jsonl)
configs = {"default_hproj_0.6_0.85": {}, "hproj_0.7": {"hfracs": [0.7]}, "hproj_after_launch": {"hfracs": []}, "hproj_0.6": {"hfracs": [0.6]}, "hproj_0.8": {"hfracs": [0.8]},
           "hproj_0.5_0.75_0.9": {"hfracs": [0.5, 0.75, 0.9]},
           "hproj_0.7_chunks_0.25_0.5_0.75": {"fracs": [0.25, 0.5, 0.75]},
           "g_fp32": {"g_half": False}, "bf16_conversion_passes": {"twins": False}, "timing_only_no_bf16_conversions": {"skip_bf16_conv": True}, "g_fp16_head0.25": {"fracs": [0.25, 0.5, 0.78]},
           # BPTT release points of the early-rows schedule (one launch publishing its progress)
           "early_0.67_0.85_0.95": {"early": [0.67, 0.85, 0.95]}, "early_0.6_0.8_0.9_0.96": {"early": [0.6, 0.8, 0.9, 0.96]},
           "early_0.7_0.9": {"early": [0.7, 0.9]}, "early_0.67_0.85_0.93_0.98": {"early": [0.67, 0.85, 0.93, 0.98]}}
ROUNDS = 7
for a in sys.argv[1:]:
    if a.startswith("--rounds="):
        ROUNDS = int(a.split("=")[1])
names = [a for a in sys.argv[1:] if not a.startswith("--")]
if names:
    configs = {k: v for k, v in configs.items() if k in names}
samples = {k: [] for k in configs}
for rnd in range(ROUNDS):                      # interleaved rounds: the power-capped clock drifts by more than the effects compared
    for k, cfg in configs.items():
        apply(cfg)
        samples[k].append(timed(6))
for k, v in samples.items():
    print(json.dumps({"config": k, "settings": configs[k], "ms_per_step_median": round(statistics.median(v), 3), "min": round(min(v), 3),
                      "max": round(max(v), 3), "rounds": len(v), "device_error": L.lcb_device_error(0)}), flush=True)
