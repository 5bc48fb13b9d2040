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
def _oracle_step(oracle, cfg, p, state, x, lens, y):
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ctc, total, _ = oracle.training_loss(pr, cfg, x, lens, y, l2_decay_weight=1e-5)
    total.backward()
    clipped, _ = oracle.clip_by_global_norm({k: v.grad for k, v in pr.items()}, 5.0)
    return oracle.adam_step({k: v.detach() for k, v in pr.items()}, clipped, state, 4e-4)


def _oracle_cfg(oracle, w):
    return oracle.OracleConfig(input_dim=w["D"], num_layers=w["num_layers"], num_neurons=w["H"], num_projects=w["P"],
                               num_targets=w["V"], use_peepholes=True, num_experts=w["K"], moe_temp=10.0)


def run_reference(args, w, emit):
    """The reference's own CPU path: TF 1.8 cannot be installed offline, so (north_star fallback) the
    PyTorch-CPU transcription of the same ops in oracle/ is timed on the host cores, all threads.

    Sample: the workload's full batch width (the per-step matmuls of the reference's while-loop are [B, .] x [., 4H]; fewer
    utterances would leave the host cores idle) x a PREFIX of the frames.  The prefix is as long as the run-time budget allows
    (<= 300 frames): one probe step on 60 frames gives the cost per frame, and the per-frame cost of the per-time-step cell does not
    depend on the prefix length, so frames/s of the prefix is the full-length figure; only the CTC lattice (a few % of the CPU
    step) scales differently."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = _oracle_cfg(oracle, w)
    Bs = min(w["B"], 64)
    p = oracle.init_params(cfg, seed=0, dtype=torch.float32)
    # probe: one step on a 60-frame prefix (also pages the libraries in)
    wp = dict(w); wp["B"], wp["T"] = Bs, min(w["T"], 60)
    xp, lp, yp = synth_batch(wp, 777)
    t0 = time.perf_counter()
    _oracle_step(oracle, cfg, p, {}, xp, lp, yp)
    per_frame = (time.perf_counter() - t0) / wp["T"]
    budget_s = 240.0
    Ts = int(budget_s / ((args.steps + args.warmup) * per_frame))
    Ts = max(min(Ts, 300, w["T"]), min(w["T"], 60)) // 20 * 20 or min(w["T"], 60)
    ws = dict(w); ws["B"], ws["T"] = Bs, Ts
    x, lens, y = synth_batch(ws, 777)
    state = {}
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        p = _oracle_step(oracle, cfg, p, state, x, lens, y)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    frames = int(lens.sum())
    val = frames / (ms / 1e3)
    sample = "%d utts x %d-frame prefix of workload %s per step (oracle port, torch CPU fp32, per-time-step cell)" % (Bs, Ts, args.workload)
    line = {"metric": "train_frames_per_sec", "value": val, "unit": "frames/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "sample": sample, "same_config": Ts == w["T"],
                       "same_config_note": "same model, batch width, optimizer and metric; frames per utterance cut to a %d-frame "
                                           "prefix so that %d CPU steps end within minutes (frames/s of the per-time-step cell does "
                                           "not depend on the prefix length)" % (Ts, args.steps + args.warmup)},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def cpu_baseline_sample(w, budget_s=25.0):
    """Oracle port timed on the host cores on a bounded sample (rank 0, N=1 only): full batch width x a 150-frame prefix."""
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = _oracle_cfg(oracle, w)
    ws = dict(w); ws["B"], ws["T"] = min(w["B"], 64), min(w["T"], 150)      # full batch width, a prefix of the frames (see run_reference)
    x, lens, y = synth_batch(ws, 777)
    p = oracle.init_params(cfg, seed=0, dtype=torch.float32)
    t_tot, frames, n = 0.0, 0, 0
    while t_tot < budget_s and n < 4:
        t0 = time.perf_counter()
        _oracle_step(oracle, cfg, p, {}, x, lens, y)
        dt = time.perf_counter() - t0
        if n > 0 or dt > budget_s / 2:          # first pass doubles as warm-up unless it is already long
            t_tot += dt; frames += int(lens.sum())
        n += 1
    if frames == 0:
        t_tot, frames = dt, int(lens.sum())
    return {"value": frames / t_tot, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "%d utts x %d-frame prefix of the workload, full training step, torch CPU fp32 oracle" % (ws["B"], ws["T"])}


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
    ap.add_argument("--no-ctc", action="store_true")
    ap.add_argument("--no-flow-control", action="store_true",
                    help="forward recurrence as consecutive range launches instead of one launch that waits in-kernel for the "
                         "projection chunks (needed when kernels cannot run concurrently, e.g. under ncu; detected automatically "
                         "from CUDA_INJECTION64_PATH / CUDA_LAUNCH_BLOCKING)")
    ap.add_argument("--fwd-hproj-frac", type=float, default=None,
                    help="A/B: share of the forward scan whose output projection runs beside the recurrence (0: all of it after the launch)")
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
    if args.no_flow_control:
        model.enc.fwd_flow_control = False
    if args.fwd_hproj_frac is not None:
        model.enc.fwd_hproj_fracs = [args.fwd_hproj_frac] if args.fwd_hproj_frac > 0 else []
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
        # (the lengths are also handed over on the host, as every batch assembler has them: the forward recurrence then skips the
        # scan steps in which a whole 16-utterance group is past its last frame)
        loss_sum, _ = model.loss_and_grad(x, lens, y, bucket_ready=reducer.bucket_ready, check_labels=False, seq_len_host=lens_h)
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

    # ---- the metric's second half: CTC loss+grad utts/s (C4 points), rank 0 ----
    ctc_rec = ctc_microbench(device) if (rank == 0 and not args.no_ctc) else None

    # ---- N > 1: the reduced gradient equals the 1-GPU gradient of the concatenated batch, and the strong-scaling figure ----
    dp = dp_check_and_strong_scaling(model, reducer, w, args, rank, world, device) if world > 1 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {"metric": "train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f16 fwd / bf16 grad operands, f32 accumulate+state+master", "data": "synthetic",
           "config": {"workload": w["desc"], "name": args.workload, "per_gpu_batch": w["B"], "frames_per_step_global": frames_global,
                      "optimizer": "adam", "keep_prob": args.keep_prob, "l2": 1e-5, "clip_norm": 5.0,
                      "l2_cache": "inputs_exceed_l2 (per-step activations >> 126 MB)", "parallelism": "dp%d" % world,
                      "fwd_flow_control": bool(model.enc.fwd_flow_control), "fwd_hproj_fracs": list(model.enc.fwd_hproj_fracs), "preactivation_dtype": "f16" if model.enc.g_half else "f32", "fwd_rec_sms": int(L.lcb_lstm_rec_grid(w["B"], model.cfg.Hp, 2, 0)),
                      "final_loss": last_loss, "device_error": dev_err},
           "clocks": clocks, "gpu_launches": launches, "e2e": e2e}
    if kt is not None:
        out["roofline"] = kt["roofline"]
        out["rooflines_other"] = kt["others"]
        out["kernels"] = kt["kernels"]
    if ctc_rec is not None:
        out["ctc"] = ctc_rec
    if dp is not None:
        out["dp_check"] = dp["dp_check"]
        out["strong_scaling"] = dp["strong_scaling"]
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
    nsm = L.lcb_device_sm_count()
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
                cap = k.get("max_ctas")                     # the persistent-grid cap this launch ran under (per-call argument)
                cap = gemm_mod.current_cap() if cap is None else cap
                gemm_calls.append((cls, 2.0 * M_ * N_ * K_, s, e, (min(int(cap), nsm) if cap and cap > 0 else nsm) / float(nsm)))
            return r
        return wrap

    orig = {"gemm_b": blstm_mod.gemm, "gemm_m": model_mod.gemm, "fwd": L.lcb_lstm_rec_fwd, "bwd": L.lcb_lstm_rec_bwd,
            "out": L.lcb_output_fwd, "ctc": model_mod.ctc_loss_grad, "mosb": L.lcb_mos_bwd_dz, "opt": L.lcb_optimizer_step}

    class Proxy:
        def __init__(self, lib):
            self._lib = lib
        def __getattr__(self, k):
            f = getattr(self._lib, k)
            m = {"lcb_lstm_rec_fwd": "lstm_rec_fwd", "lcb_lstm_rec_fwd_range": "lstm_rec_fwd", "lcb_lstm_rec_fwd_range_hl": "lstm_rec_fwd", "lcb_lstm_rec_fwd_range_pg": "lstm_rec_fwd", "lcb_lstm_rec_bwd": "lstm_rec_bwd", "lcb_lstm_rec_bwd_range": "lstm_rec_bwd", "lcb_lstm_rec_bwd_range_pg": "lstm_rec_bwd", "lcb_output_fwd": "output_fwd",
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
            model.loss_and_grad(x, lens, y, check_labels=False, seq_len_host=lens.cpu())
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
    B, T, H, Ltot = w["B"], w["T"], w["H"], w["num_layers"]
    Hp = (H + 63) // 64 * 64
    traffic = {}
    try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        traffic = tr.get(w.get("desc", ""), {})
    except Exception:
        pass

    def ms_of(*names):
        return sum(kernels[n]["ms_total"] for n in names if n in kernels)

    def launches_of(*names):
        return sum(kernels[n]["launches"] for n in names if n in kernels)

    roofs = {}
    # (1) the recurrence family: forward recurrence + BPTT (cluster-persistent kernels).  Neither tensor- nor HBM-bound: 2 x T
    # dependent steps per layer on 64 of the SMs -- the figure of merit is microseconds per time step against the floor of the
    # decomposition; achieved / peak are kept in the contract's units (algorithmic flops of the recurrent product and
    # its dgrad over VALID frames).
    ms_f, ms_b = ms_of("lstm_rec_fwd"), ms_of("lstm_rec_bwd")
    if ms_f + ms_b > 0:
        flops = 2.0 * recurrent_flops_per_frame(w) * frames          # forward product + its dgrad in BPTT
        ach = flops / ((ms_f + ms_b) * 1e-3) / 1e12
        # algorithmic HBM bytes per time step and utterance: forward reads G (8Hp f32) and writes m (2Hp f16), gates (2Hp x 8 B),
        # c (2Hp f32); BPTT reads gates, c (twice: c_t and c_{t-1}), dM (2Hp f32) and writes dG (8Hp bf16)
        by_f = (8 * Hp * 4 + 2 * Hp * (2 + 8 + 4)) * float(frames) * Ltot
        by_b = (2 * Hp * (8 + 4 + 4 + 4) + 8 * Hp * 2) * float(frames) * Ltot
        tr_f, tr_b = traffic.get("lstm_rec_fwd"), traffic.get("lstm_rec_bwd")
        clk = 1.965e9
        roofs["lstm_rec"] = {
            "kernel": "lstm_rec (lstm_rec_fwd2 + lstm_rec_bwd3: forward recurrence and BPTT, all layers, both directions)",
            "bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
            "traffic": ({"lstm_rec_fwd": tr_f["dram_bytes_per_launch"], "lstm_rec_bwd": tr_b["dram_bytes_per_launch"],
                         "source": tr_f.get("source", "")} if (tr_f and tr_b) else None),
            "peak_source": how + " sustained (kernels timed inside a long step)",
            "latency_bound": True,
            "us_per_time_step": {"fwd": ms_f * 1e3 / (Ltot * T), "bwd": ms_b * 1e3 / (Ltot * T)},
            "floor_us_per_time_step": {"fwd": 1800.0 / clk * 1e6, "bwd": 2400.0 / clk * 1e6,
                                       "how": "cycles of the serial chain this decomposition cannot shed, at 1.965 GHz: one weight pass "
                                              "out of tensor memory (128 KB per CTA and step, ~250 B/clk measured = 526) + two L2 trips "
                                              "of the m_t exchange (bulk store 500 + multicast 670, probes) + the dependent MUFU chain "
                                              "of the gate math (~100); BPTT adds the 4-way reduce-scatter over DSMEM (~600)"},
            "algorithmic_hbm_bytes": {"fwd": by_f, "bwd": by_b},
            "hbm_gbs": {"fwd": by_f / (ms_f * 1e-3) / 1e9 if ms_f else None, "bwd": by_b / (ms_b * 1e-3) / 1e9 if ms_b else None,
                        "peak": hbm},
            "launches": launches_of("lstm_rec_fwd", "lstm_rec_bwd"), "ms_total": ms_f + ms_b,
            "note": "share of the step: see kernels[*].share; achieved = 2 x recurrent_flops_per_frame x valid frames / summed launch time"}
    # (2) bulk GEMM family
    if "gemm" in kernels:
        flops = gemm_family_flops_per_frame(w) * frames
        ms1 = kernels["gemm"]["ms_total"]
        ach = flops / (ms1 * 1e-3) / 1e12
        roofs["gemm"] = {"kernel": "gemm16 (all bulk GEMMs of the step: projections, dgrad, wgrad, output layer)", "bound": "tensor",
                         "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
                         "traffic": (traffic.get("gemm") or {}).get("dram_bytes_per_launch"),
                         "peak_source": how + " sustained (MEASURED_PEAKS.json, unscaled: the whole chip's cuBLAS bf16 peak)",
                         "launches": kernels["gemm"]["launches"], "ms_total": ms1,
                         "note": "algorithmic flops over VALID frames / summed GEMM launch time; many launches run on a capped "
                                 "persistent grid beside the recurrence clusters (sm_share_* keys), so frac understates what the "
                                 "kernel does on the SMs it is given"}
    # (3) CTC (in-step shape)
    if "ctc_loss_grad" in kernels:
        bytes_ = 8.0 * T * B * w["V"]
        k_ = kernels["ctc_loss_grad"]
        ach = bytes_ / (k_["ms_total"] / k_["launches"] * 1e-3) / 1e9
        roofs["ctc_loss_grad"] = {"kernel": "ctc_loss_grad (in-step shape B=%d T=%d V=%d)" % (B, T, w["V"]), "bound": "hbm",
                                  "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                                  "traffic": (traffic.get("ctc_loss_grad") or {}).get("dram_bytes_per_launch"),
                                  "peak_source": how, "ms_per_launch": k_["ms_total"] / k_["launches"]}
    # per-class GEMM throughput of the same instrumented step (launched flops incl. padded frames / launch time); launches of
    # >= 20 GFLOP only, so the tiny weight-folding GEMMs do not blur the classes.
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
        c_["frac"] = c_["tflops"] / tf_sus
        c_["mean_sm_share"] = c_["sm_ms"] / c_["ms_total"] if c_["ms_total"] > 0 else 1.0
        c_["frac_of_sustained_peak_of_sms_granted"] = c_["tflops"] / (tf_sus * c_["mean_sm_share"])
        del c_["flops"], c_["sm_ms"]
    if "gemm" in roofs and all_ms > 0:
        roofs["gemm"]["gemm_classes"] = classes
        roofs["gemm"]["sm_share_mean"] = share_ms / all_ms
        roofs["gemm"]["sm_share_frac"] = roofs["gemm"]["achieved"] / (tf_sus * share_ms / all_ms)
    # dominant family by summed launch time (forward recurrence and BPTT count as ONE family)
    fam_ms = {"lstm_rec": ms_f + ms_b, "gemm": ms_of("gemm"), "ctc_loss_grad": ms_of("ctc_loss_grad")}
    dom = max((k for k in fam_ms if k in roofs), key=lambda k: fam_ms[k])
    roof = roofs.pop(dom)
    roof["dominant_by"] = "summed launch time of the instrumented step (ms): " + json.dumps({k: round(v, 3) for k, v in fam_ms.items()})
    return {"kernels": kernels, "roofline": roof, "others": roofs}


def ctc_microbench(device, B=256, T=700, Llab=100, Vs=(72, 5000), iters=5):
    """CTC loss+grad utts/s (the second half of BASELINE.json's metric) on two points of the C4 sweep, timed with CUDA events.
    Inputs are rotated over enough distinct buffers that consecutive launches never find their logits in the 126 MB L2."""
    from lstm_ctc_b200.ctc import ctc_loss_grad
    hbm = _peaks()[0]
    out = []
    for V in Vs:
        g = torch.Generator().manual_seed(0)
        bytes_one = 4.0 * B * T * V
        nbuf = 1 if bytes_one > 512e6 else int(min(8, max(2, 300e6 // bytes_one + 1)))
        xs = [(torch.randn(B, T, V, generator=g) * 3).to(device) for _ in range(nbuf)]
        sl = torch.randint(int(0.8 * T), T + 1, (B,), generator=g).to(torch.int32).to(device)
        lab = torch.randint(0, V - 1, (B, Llab), generator=g).to(device)
        for k in range(3):
            ctc_loss_grad(xs[k % nbuf], lab, sl, check_labels=False)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for k in range(iters):
            ctc_loss_grad(xs[k % nbuf], lab, sl, check_labels=False)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        gb = 8.0 * T * B * V / 1e9
        out.append({"T": T, "L": Llab, "V": V, "B": B, "ms": ms, "utts_per_s": B / ms * 1e3, "gbs": gb / ms * 1e3,
                    "peak_gbs": hbm, "frac": gb / ms * 1e3 / hbm, "input_buffers_rotated": nbuf})
        del xs
    return {"metric": "ctc_loss_grad_utts_per_sec", "unit": "utts/s", "algorithmic_bytes": "8*T*B*V", "points": out}


def dp_check_and_strong_scaling(model, reducer, w, args, rank, world, device):
    """(a) dp_check: the SUM all-reduced gradient of a global batch of w['B'] utterances dealt round-robin over the ranks equals the
    gradient rank 0 computes alone on the whole batch (summed loss graph.py:116, global norm graph.py:190) -- with keep_prob 1
    (dropout masks are indexed by the element's position in the local batch, so they differ between the two runs by design).
    (b) strong scaling: the same global batch, training steps timed like the headline (max over ranks)."""
    import torch.distributed as dist
    xg, lg, yg = synth_batch(w, 777)
    idx = list(range(rank, w["B"], world))
    xs, ls, ys = xg[idx].to(device), lg[idx].to(device), yg[idx].to(device)
    keep = model.cfg.keep_prob
    model.cfg.keep_prob = 1.0
    reducer.broadcast_weights()

    def compare(Tc):
        """reduced gradient of the sharded batch vs rank 0's gradient of the whole batch, frames per utterance cut to Tc"""
        lc = torch.clamp(lg, max=Tc)
        yc = yg.clone()
        for b_ in range(yc.shape[0]):                      # keep the labels feasible for the cut utterance
            yc[b_, max(1, int(lc[b_]) // 8):] = -1
        xs_, ls_, ys_ = xg[idx, :Tc].contiguous().to(device), lc[idx].to(device), yc[idx].to(device)
        reducer.begin_step()
        loss_s, _ = model.loss_and_grad(xs_, ls_, ys_, bucket_ready=reducer.bucket_ready, check_labels=False)
        reducer.finish()
        lsum = loss_s.double().reshape(1).clone()
        dist.all_reduce(lsum)
        torch.cuda.synchronize()
        out = None
        if rank == 0:
            g_dp = model.params.gflat.double().clone()
            loss_1, _ = model.loss_and_grad(xg[:, :Tc].contiguous().to(device), lc.to(device), yc.to(device), check_labels=False)
            g_1 = model.params.gflat.double()
            n1, ndp = float(g_1.norm()), float(g_dp.norm())
            out = {"frames_per_utt": Tc, "grad_rel_err": float((g_dp - g_1).norm()) / max(n1, 1e-30), "global_norm_dp": ndp,
                   "global_norm_1gpu": n1, "loss_sum_dp": float(lsum), "loss_sum_1gpu": float(loss_1.double())}
        dist.barrier()
        torch.cuda.synchronize()
        return out

    # Full length: a randomly initialised 5-layer stack with forget bias 5 has a gradient norm of ~1e11 at T = 1500: BPTT amplifies
    # any perturbation of d loss / d logits by that much.  What differs between the sharded and the unsharded run is only the ORDER
    # of fp32 additions (the red.global.add of the CTC gammas inside an utterance, the K loops / split-K of the weight-gradient
    # GEMMs; forward activations are bit-identical per utterance), and that order noise shows at 1e-5 .. 4e-4 of the norm, varying
    # from run to run (measured: 7e-6 .. 4e-4 at 2 ranks, 1e-4 .. 3e-4 at 4).  The 256-frame cut
    # of the same batch (gradient norm ~1e7, no such amplification) pins the exchange itself: 4e-6.
    full, cut = compare(w["T"]), compare(min(256, w["T"]))
    res = None
    if rank == 0:
        tol = {"grad_rel_err_full_length": 2e-3, "grad_rel_err_256_frames": 2e-5, "norm_rel": 1e-5, "loss_rel": 1e-5}

        def fine(r, gt):
            return bool(r["grad_rel_err"] < gt and abs(r["global_norm_dp"] - r["global_norm_1gpu"]) <= max(tol["norm_rel"], gt) * r["global_norm_1gpu"]
                        and abs(r["loss_sum_dp"] - r["loss_sum_1gpu"]) <= tol["loss_rel"] * abs(r["loss_sum_1gpu"]))
        res = {"global_batch": w["B"], "ranks": world, "keep_prob": 1.0, "full_length": full, "cut_256_frames": cut, "tolerance": tol,
               "grad_rel_err": full["grad_rel_err"],
               "ok": fine(full, tol["grad_rel_err_full_length"]) and fine(cut, tol["grad_rel_err_256_frames"])}
    model.cfg.keep_prob = keep
    dist.barrier()
    torch.cuda.synchronize()

    def step():
        reducer.begin_step()
        model.loss_and_grad(xs, ls, ys, bucket_ready=reducer.bucket_ready, check_labels=False)
        reducer.finish()
        model.optimizer_step("adam", 4e-4, clip_norm=5.0, l2_decay_weight=1e-5)

    for _ in range(max(args.warmup, 3)):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    frames = float(lg.sum())
    strong = {"scaling": "strong", "global_batch": w["B"], "per_gpu_batch": len(idx), "ms_per_step": ms_step,
              "value": frames / (ms_step / 1e3), "unit": "frames/s", "n_gpus": world}
    return {"dp_check": res, "strong_scaling": strong}


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
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 32, "api": "create_graph_for_training_ctc + Session.run",
            "d2h": "4 x f64 per step (CTC loss sum, token count, label-smoothing term, spare), copied right behind the CTC kernels",
            "last_eval_loss": vals["eval_loss"]}


if __name__ == "__main__":
    main()
