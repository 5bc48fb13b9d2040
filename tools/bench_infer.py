"""C5 (BASELINE.json configs[4]): forward-only BiLSTM-MoS posterior path of bin/nnet-forward.py -- WSJ-shape model
(4 x BiLSTM 320/dir, K=8 mixture output, V=72), B=512 utterances x ~700 frames, logits -> log softmax - log prior.
Prints one JSON line: frames/s and ms/batch at B=512, and the B=1 latency (the reference's own mode: one utterance per
sess.run, nnet-forward.py:81-83), device-resident and with the host->device copy of the features + device->host copy of
the posteriors inside the timed region."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lstm_ctc_b200.decode import softmax_rows  # noqa: E402
from lstm_ctc_b200.model import AcousticModel  # noqa: E402


def run(B, T, iters, e2e):
    w = dict(bench.WORKLOADS["c2"]); w["B"], w["T"] = B, T
    dev = torch.device("cuda:0")
    cfg = bench.nnet_config(w, 1.0); cfg["is_training"] = False
    model = AcousticModel(cfg, dev, seed=1234)
    x_h, lens_h, _ = bench.synth_batch(w, 777)
    x_h = x_h.pin_memory()
    prior = torch.log_softmax(torch.randn(w["V"]), 0).to(dev)
    x, lens = x_h.to(dev), lens_h.to(dev)
    out_h = torch.empty(B, T, w["V"], dtype=torch.float32).pin_memory()

    def step():
        xi = x_h.to(dev, non_blocking=True) if e2e else x
        logits = model.forward_logits(xi, lens, training=False)
        post = softmax_rows(logits, 1.0, apply_log=True, log_prior=prior)
        if e2e:
            out_h.copy_(post, non_blocking=True)
        return post

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(iters):
        step()
    e.record(); torch.cuda.synchronize()
    ms = max(s.elapsed_time(e), (time.perf_counter() - t0) * 1e3 if e2e else 0.0) / iters
    return ms, float(lens_h.sum())


if __name__ == "__main__":
    res = {"workload": "C5: WSJ-shape BiLSTM-MoS forward + log-softmax - prior (nnet-forward.py posterior path)", "dtype": "f16 operands, f32 accumulate/state"}
    for B, T, it in ((512, 700, 5), (64, 700, 10), (1, 700, 20)):
        for e2e in (False, True):
            ms, frames = run(B, T, it, e2e)
            res["B%d_%s" % (B, "e2e" if e2e else "device")] = {"ms_per_batch": ms, "frames_per_s": frames / ms * 1e3, "utts_per_s": B / ms * 1e3}
        torch.cuda.empty_cache()
    print(json.dumps(res))
