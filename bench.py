#!/usr/bin/env python
"""bench.py -- train frames/sec of the BiLSTM-MoS-CTC hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                    (the reference's CPU path: oracle port, host cores)

A "step" is one full training step of the reference (forward, CTC loss+grad, backward, gradient all-reduce,
L2 + global-norm clip + Adam) on one synthetic minibatch per GPU.  Workload (default `c3`): the
LibriSpeech-shape configuration the metric is quoted on: 5 x BiLSTM 512 cells/dir (num_projects 512,
peepholes), 120-dim input, mixture output K=8 tau=10, V=72, dropout keep-prob 0.9, 64 utterances x 1500 frames
per GPU (weak scaling), lengths ~ U{0.8T..T}, labels ~ len/8.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (layers, H, P, D, V, K, B per GPU, T)
    "c3": dict(num_layers=5, H=512, P=512, D=120, V=72, K=8, B=64, T=1500,
               desc="LibriSpeech-shape BiLSTM-MoS-CTC: 5xBiLSTM 512/dir, K=8, B=64 utts x ~1500 frames per GPU"),
    "c2": dict(num_layers=4, H=320, P=320, D=120, V=72, K=8, B=64, T=700,
               desc="WSJ-phone high-rank model: 4xBiLSTM 320/dir, K=8, B=64 x ~700 frames"),
    "c1": dict(num_layers=4, H=320, P=320, D=120, V=72, K=0, B=16, T=700,
               desc="WSJ-phone BiLSTM-CTC: 4xBiLSTM 320/dir, affine output, B=16 x ~700 frames"),
    "tiny": dict(num_layers=2, H=128, P=128, D=40, V=30, K=4, B=16, T=100, desc="smoke-size"),
}


def nnet_config(w, keep=0.9):
    return {"nnet_type": "blstm", "input_dim": w["D"], "left_context": 0, "right_context": 0, "subsample": 0,
            "num_layers": w["num_layers"], "num_neurons": w["H"], "num_projects": w["P"], "num_targets": w["V"],
            "use_peepholes": True, "num_experts": w["K"], "moe_temp": 10.0, "dropout_rate": keep, "is_training": True}


def train_flops_per_frame(w):
    """GEMM flops per valid frame, forward; training = 3x."""
    H, P, D, V, K, L = w["H"], w["P"], w["D"], w["V"], w["K"], w["num_layers"]
    f = 0.0
    for i in range(L):
        din = D if i == 0 else 2 * P
        f += 2 * (2 * (din + P) * 4 * H + 2 * H * P)
    f += 2 * 2 * P * (K * (V + 1) if K > 0 else V)
    return 3.0 * f


def recurrent_flops_per_frame(w):
    """The serial h_{t-1} * W_hh part of the forward flops (both directions, all layers) -- done inside the recurrence
    kernels (forward) and the BPTT kernels (its dgrad); only its weight gradient is a bulk GEMM."""
    return w["num_layers"] * 2 * 2.0 * w["P"] * 4 * w["H"]


def gemm_family_flops_per_frame(w):
    """Algorithmic flops the bulk GEMM launches of one training step carry per valid frame: everything of SURVEY 8(d)'s
    3 x F_fwd except the recurrent product's forward and dgrad (which run in lstm_rec_fwd / lstm_rec_bwd)."""
    return train_flops_per_frame(w) - 2.0 * recurrent_flops_per_frame(w)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(w, seed, device=None):
    g = np.random.RandomState(seed)
    B, T, D, V = w["B"], w["T"], w["D"], w["V"]
    lens = np.sort(g.randint(int(np.ceil(0.8 * T)), T + 1, size=B)).astype(np.int32)
    lens[-1] = T
    Lmax = T // 8
    x = torch.zeros(B, T, D, dtype=torch.float32)
    y = torch.full((B, Lmax), -1, dtype=torch.int64)
    for b in range(B):
        n = int(lens[b])
        x[b, :n] = torch.from_numpy(g.standard_normal((n, D)).astype(np.float32))
        L = max(1, n // 8)
        y[b, :L] = torch.from_numpy(g.randint(0, V - 1, size=L))
    return x, torch.from_numpy(lens), y


# ------------------------------------------------------------------------------------------------
def run_reference(args, w, emit):
    """The reference's own CPU path: TF 1.8 cannot be installed offline, so (north_star fallback) the
    PyTorch-CPU transcription of the same ops in oracle/ is timed on the host cores, all threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = oracle.OracleConfig(input_dim=w["D"], num_layers=w["num_layers"], num_neurons=w["H"], num_projects=w["P"],
                              num_targets=w["V"], use_peepholes=True, num_experts=w["K"], moe_temp=10.0)
    # bounded sample of the same workload: the workload's batch width (the per-step matmuls of the reference's while-loop are
    # [B, .] x [., 4H]: two utterances would leave the host cores idle -- 86 vs 670 frames/s on 8 cores), a prefix of the frames
    Bs, Ts = min(w["B"], 64), min(w["T"], 60)
    ws = dict(w); ws["B"], ws["T"] = Bs, Ts
    x, lens, y = synth_batch(ws, 777)
    p = oracle.init_params(cfg, seed=0, dtype=torch.float32)
    state = {}
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        ctc, total, _ = oracle.training_loss(pr, cfg, x, lens, y, l2_decay_weight=1e-5)
        total.backward()
        clipped, _ = oracle.clip_by_global_norm({k: v.grad for k, v in pr.items()}, 5.0)
        p = oracle.adam_step({k: v.detach() for k, v in pr.items()}, clipped, state, 4e-4)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    frames = int(lens.sum())
    val = frames / (ms / 1e3)
    sample = "%d utts x %d frames of workload %s per step (oracle port, torch CPU fp32, per-time-step cell)" % (Bs, Ts, args.workload)
    line = {"metric": "train_frames_per_sec", "value": val, "unit": "frames/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "sample": sample},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def cpu_baseline_sample(w, budget_s=20.0):
    """Oracle port timed on the host cores on a bounded sample (rank 0, N=1 only)."""
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = oracle.OracleConfig(input_dim=w["D"], num_layers=w["num_layers"], num_neurons=w["H"], num_projects=w["P"],
                              num_targets=w["V"], use_peepholes=True, num_experts=w["K"], moe_temp=10.0)
    ws = dict(w); ws["B"], ws["T"] = min(w["B"], 64), min(w["T"], 60)      # full batch width, a prefix of the frames (see run_reference)
    x, lens, y = synth_batch(ws, 777)
    p = oracle.init_params(cfg, seed=0, dtype=torch.float32)
    t_tot, frames, n = 0.0, 0, 0
    while t_tot < budget_s and n < 3:
        t0 = time.perf_counter()
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        ctc, total, _ = oracle.training_loss(pr, cfg, x, lens, y, l2_decay_weight=1e-5)
        total.backward()
        clipped, _ = oracle.clip_by_global_norm({k: v.grad for k, v in pr.items()}, 5.0)
        oracle.adam_step({k: v.detach() for k, v in pr.items()}, clipped, {}, 4e-4)
        dt = time.perf_counter() - t0
        if n > 0 or dt > budget_s / 2:          # first pass doubles as warm-up unless it is already long
            t_tot += dt; frames += int(lens.sum())
        n += 1
    if frames == 0:
        t_tot, frames = dt, int(lens.sum())
    return {"value": frames / t_tot, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "%d utts x %d frames of the workload, full training step, torch CPU fp32 oracle" % (ws["B"], ws["T"])}


# ------------------------------------------------------------------------------------------------
def main():
    # the driver expects exactly ONE JSON line on stdout: route everything else (NCCL banners, warnings) to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--keep-prob", type=float, default=0.9)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, emit)
        return
    args.warmup = max(args.warmup, 3)

    import lstm_ctc_b200 as nnet
    from lstm_ctc_b200 import _lib, dist as lcb_dist
    from lstm_ctc_b200.model import AcousticModel
    import torch.distributed as dist

    rank, world, device = lcb_dist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    L = _lib.lib()
    cfg = nnet_config(w, args.keep_prob)
    model = AcousticModel(cfg, device, seed=1234)
    reducer = lcb_dist.GradientAllReducer(model.params)
    reducer.broadcast_weights()
    x_h, lens_h, y_h = synth_batch(w, 777 + rank)
    x, lens, y = x_h.to(device), lens_h.to(device), y_h.to(device)
    frames_local = int(lens_h.sum())
    frames_t = torch.tensor([float(frames_local)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(frames_t)
    frames_global = float(frames_t.item())

    rec = {"rec_fwd": [], "rec_bwd": []}

    def step(instrument=False):
        reducer.begin_step()
        loss_sum, _ = model.loss_and_grad(x, lens, y, bucket_ready=reducer.bucket_ready, check_labels=False)
        reducer.finish()
        model.optimizer_step("adam", 4e-4, clip_norm=5.0, l2_decay_weight=1e-5)
        return loss_sum

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (the `value`) ----
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(device.index or 0)
    if rank == 0:
        sampler.start()
    L.lcb_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss_sum = step()
    e1.record()
    barrier()
    launches = int(L.lcb_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = frames_global / (ms_per_step / 1e3)
    last_loss = float(loss_sum.item())
    dev_err = L.lcb_device_error(0)

    # ---- per-kernel durations, measured live with CUDA events on the launching stream ----
    kt = kernel_breakdown(model, x, lens, y, w, frames_local) if rank == 0 else None

    # ---- e2e through the public API with HOST buffers (H2D of inputs + D2H of the loss every step) ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(nnet, cfg, w, x_h, lens_h, y_h, args, world, device, frames_global)

    if rank != 0:
        return
    out = {"metric": "train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f16 fwd / bf16 grad operands, f32 accumulate+state+master", "data": "synthetic",
           "config": {"workload": w["desc"], "name": args.workload, "per_gpu_batch": w["B"], "frames_per_step_global": frames_global,
                      "optimizer": "adam", "keep_prob": args.keep_prob, "l2": 1e-5, "clip_norm": 5.0,
                      "l2_cache": "inputs_exceed_l2 (per-step activations >> 126 MB)", "parallelism": "dp%d" % world,
                      "final_loss": last_loss, "device_error": dev_err},
           "clocks": clocks, "gpu_launches": launches, "e2e": e2e}
    if kt is not None:
        out["roofline"] = kt["roofline"]
        out["kernels"] = kt["kernels"]
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample(w)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("hbm_gbs", 6650.0), p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def kernel_breakdown(model, x, lens, y, w, frames):
    """Time each kernel family of one training step with CUDA events on the launching stream (one extra
    instrumented step after the timed region; the un-instrumented timed region above is the headline)."""
    from lstm_ctc_b200 import _lib, gemm as gemm_mod, blstm as blstm_mod, model as model_mod, ctc as ctc_mod
    L = _lib.lib()
    spans = {}

    gemm_calls = []                 # (class, flops, start, end) of every bulk GEMM launch of the instrumented step

    def timed(name, fn):
        def wrap(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            spans.setdefault(name, []).append((s, e))
            if name == "gemm":
                A, Bm = a[0], a[1]
                al = a[2] if len(a) > 2 else k.get("a_layout", 0)
                bl = a[3] if len(a) > 3 else k.get("b_layout", 0)
                M_, K_ = (A.shape[0], A.shape[1]) if al == 0 else (A.shape[1], A.shape[0])
                N_ = Bm.shape[0] if bl == 0 else Bm.shape[1]
                cls = {(0, 0): "forward (X*W^T: projections, h-projection, output recompute)", (0, 1): "dgrad (dG*W)",
                       (1, 1): "wgrad (X^T*dG, K = frames)"}.get((al, bl), "other")
                cap = L.lcb_gemm_set_max_ctas(148)          # read the persistent-grid cap this launch ran under (host-side setting)
                L.lcb_gemm_set_max_ctas(cap)
                gemm_calls.append((cls, 2.0 * M_ * N_ * K_, s, e, min(max(int(cap), 1), 148) / 148.0))
            return r
        return wrap

    orig = {"gemm_b": blstm_mod.gemm, "gemm_m": model_mod.gemm, "fwd": L.lcb_lstm_rec_fwd, "bwd": L.lcb_lstm_rec_bwd,
            "out": L.lcb_output_fwd, "ctc": model_mod.ctc_loss_grad, "mosb": L.lcb_mos_bwd_dz, "opt": L.lcb_optimizer_step}

    class Proxy:
        def __init__(self, lib):
            self._lib = lib
        def __getattr__(self, k):
            f = getattr(self._lib, k)
            m = {"lcb_lstm_rec_fwd": "lstm_rec_fwd", "lcb_lstm_rec_fwd_range": "lstm_rec_fwd", "lcb_lstm_rec_bwd": "lstm_rec_bwd", "lcb_lstm_rec_bwd_range": "lstm_rec_bwd", "lcb_output_fwd": "output_fwd",
                 "lcb_mos_bwd_dz": "mos_bwd_dz", "lcb_optimizer_step": "optimizer"}.get(k)
            return timed(m, f) if m else f

    proxy = Proxy(L)
    old_lib = _lib.lib
    _lib.lib = lambda: proxy
    blstm_mod.gemm = timed("gemm", orig["gemm_b"])
    model_mod.gemm = timed("gemm", orig["gemm_m"])
    model_mod.ctc_loss_grad = timed("ctc_loss_grad", orig["ctc"])
    try:
        for _ in range(2):
            spans.clear()
            del gemm_calls[:]
            model.loss_and_grad(x, lens, y, check_labels=False)
            model.optimizer_step("adam", 4e-4)
            torch.cuda.synchronize()
    finally:
        _lib.lib = old_lib
        blstm_mod.gemm, model_mod.gemm, model_mod.ctc_loss_grad = orig["gemm_b"], orig["gemm_m"], orig["ctc"]
    kernels = {k: {"launches": len(v), "ms_total": sum(s.elapsed_time(e) for s, e in v)} for k, v in spans.items()}
    tot = sum(v["ms_total"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = v["ms_total"] / tot if tot else 0.0
    hbm, tf_burst, tf_sus, how = _peaks()
    # dominant kernel family by time
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"])
    B, T, H, Ltot = w["B"], w["T"], w["H"], w["num_layers"]
    Hp = (H + 63) // 64 * 64
    roof = None
    if dom in ("lstm_rec_fwd", "lstm_rec_bwd"):
        # algorithmic flops of the family per step: the folded recurrent product m_{t-1} W' (or its dgrad W' dz_t), both
        # directions, all layers, VALID frames only (padded frames are not algorithmic work)
        flops = recurrent_flops_per_frame(w) * frames
        ms_all = kernels[dom]["ms_total"]
        ach = flops / (ms_all * 1e-3) / 1e12
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
                "traffic": None, "peak_source": how + " sustained (kernel timed inside a long step)",
                "note": "serial recurrence: latency-bound, not tensor- or HBM-bound (2 x T dependent steps per layer, 64 of 148 SMs); "
                        "us_per_time_step = %.3f" % (ms_all * 1e3 / (Ltot * T))}
    elif dom == "gemm":
        flops = gemm_family_flops_per_frame(w) * frames
        ms1 = kernels[dom]["ms_total"]
        ach = flops / (ms1 * 1e-3) / 1e12
        roof = {"kernel": "gemm16 (all bulk GEMMs of the step: projections, dgrad, wgrad, output layer)", "bound": "tensor",
                "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus, "traffic": None,
                "peak_source": how + " sustained",
                "note": "algorithmic flops over VALID frames / summed GEMM launch time (weight-gradient GEMMs run on a capped grid "
                        "beside the BPTT clusters, so their launch time overlaps other kernels)"}
    else:
        bytes_ = 8.0 * T * B * w["V"]
        ms1 = kernels.get("ctc_loss_grad", {"ms_total": 1.0, "launches": 1})
        ach = bytes_ / (ms1["ms_total"] / ms1["launches"] * 1e-3) / 1e9
        roof = {"kernel": "ctc_loss_grad", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": None, "peak_source": how}
    # per-class GEMM throughput of the same instrumented step (launched flops incl. padded frames / launch time); launches of
    # >= 20 GFLOP only, so the tiny weight-folding GEMMs do not blur the classes.  The forward projections' tails and all
    # weight gradients run on a CAPPED grid (84 / 80 of 148 SMs) beside the recurrence, by design.
    classes = {}
    share_ms = 0.0                      # sum over GEMM launches of duration x fraction of the SMs the launch was granted
    all_ms = 0.0
    for cls, fl, s_, e_, share in gemm_calls:
        ms_ = s_.elapsed_time(e_)
        share_ms += ms_ * share
        all_ms += ms_
        if fl < 2e10:
            continue
        c_ = classes.setdefault(cls, {"launches": 0, "flops": 0.0, "ms_total": 0.0, "sm_ms": 0.0})
        c_["launches"] += 1
        c_["flops"] += fl
        c_["ms_total"] += ms_
        c_["sm_ms"] += ms_ * share
    for c_ in classes.values():
        c_["tflops"] = c_["flops"] / (c_["ms_total"] * 1e-3) / 1e12 if c_["ms_total"] > 0 else 0.0
        c_["mean_sm_share"] = c_["sm_ms"] / c_["ms_total"] if c_["ms_total"] > 0 else 1.0
        c_["frac_of_sustained_peak_of_sms_granted"] = c_["tflops"] / (tf_sus * c_["mean_sm_share"])
        del c_["flops"], c_["sm_ms"]
    if roof is not None and classes:
        roof["gemm_classes"] = classes
    if roof is not None and dom == "gemm" and all_ms > 0:
        # Many launches of this family run on a CAPPED persistent grid (84 / 80 of 148 SMs) beside the recurrence clusters, by
        # design; the roofline of such a launch is the peak of the SMs it was granted.  `peak` is therefore the measured
        # sustained peak x the duration-weighted mean SM share of the family's launches; the whole-chip figures stay beside it.
        mean_share = share_ms / all_ms
        roof["peak_whole_chip"] = roof["peak"]
        roof["frac_whole_chip"] = roof["frac"]
        roof["mean_sm_share"] = mean_share
        roof["peak"] = roof["peak_whole_chip"] * mean_share
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["note"] += "; peak = measured sustained peak x duration-weighted mean fraction of the 148 SMs the launches were granted"
    # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed `ncu --set full`
    # capture of this same command (profiles/r01_ncu_traffic.json; null if the capture does not cover this workload)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        ent = tr.get(w.get("desc", ""), {}).get(dom if dom != "gemm" else "gemm")
        if ent is not None:
            roof["traffic"] = ent["dram_bytes_per_launch"]
            roof["traffic_source"] = ent["source"]
    except Exception:
        pass
    return {"kernels": kernels, "roofline": roof}


def run_e2e(nnet, cfg, w, x_h, lens_h, y_h, args, world, device, frames_global):
    """Same metric through the reference-facing API: pipeline of pinned HOST tensors ->
    create_graph_for_training_ctc -> Session.run(graph nodes); every step copies its inputs H2D and reads
    the loss back D2H."""
    import torch.distributed as dist

    class OneBatchDataset:                      # yields the utterances of this rank's batch, steps+warmup times
        def __init__(self, reps):
            self.reps = reps
        def __iter__(self):
            for _ in range(self.reps):
                for b in range(x_h.shape[0]):
                    n = int(lens_h[b])
                    yield {"nnet_input": x_h[b, :n].numpy(), "nnet_target": y_h[b][y_h[b] >= 0].numpy()}

    steps, warm = args.steps, args.warmup
    init, pipeline = nnet.create_pipeline_sequence_batch(OneBatchDataset(steps + warm), w["D"], batch_size=w["B"])
    graph = nnet.create_graph_for_training_ctc(pipeline, cfg, learn_rate=4e-4, clip_norm=5.0, optimizer="adam", seed=1234)
    sess = nnet.Session()
    sess.run(init)
    nodes = {k: graph[k] for k in ("size", "train", "loss", "eval_loss", "sequence_length")}
    for _ in range(warm):
        sess.run(nodes)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        vals = sess.run(nodes)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    ms = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / steps
    h2d = x_h.numel() * 4 + lens_h.numel() * 4 + y_h.numel() * 8
    return {"value": frames_global / (ms_step / 1e3), "unit": "frames/s", "ms_per_step": ms_step,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8, "api": "create_graph_for_training_ctc + Session.run",
            "last_eval_loss": vals["eval_loss"]}


if __name__ == "__main__":
    main()
