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


def run_cli(n_utts, T, batch_sizes, tmp="/tmp/lcb_c5_cli"):
    """The same path through bin/nnet-forward.py itself (cli.nnet_forward): synthetic TFRecords on local disk -> scp -> batched
    engine -> Kaldi ark.  Reports utterances/s end to end and where the time goes (device seconds of the engine vs wall clock:
    the rest is TFRecord decoding, pinned staging and archive writing on the host)."""
    import os
    import shutil
    from lstm_ctc_b200 import cli, tfrecord as tfr
    from lstm_ctc_b200.graph import Saver
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    w = dict(bench.WORKLOADS["c2"])
    rng = np.random.RandomState(0)
    lens = rng.randint(int(0.8 * T), T + 1, size=n_utts)
    scp = os.path.join(tmp, "feats.scp")
    with open(scp, "w") as fh:
        for i, n in enumerate(lens):
            pth = os.path.join(tmp, "utt%05d.tfrecords" % i)
            tfr.write_tfrecord(pth, rng.standard_normal((n, w["D"])).astype(np.float32), None)
            fh.write("utt%05d %d %d 0 %s\n" % (i, n, w["D"], pth))
    cfg = os.path.join(tmp, "nnet.config")
    with open(cfg, "w") as fh:
        fh.write("nnet_type blstm\ninput_dim %d\nleft_context 0\nright_context 0\nsubsample 0\nnum_layers %d\nnum_neurons %d\n"
                 "num_projects %d\nnum_targets %d\nuse_peepholes true\nnum_experts %d\nmoe_temp 10.0\ndropout_rate 1.0\n"
                 % (w["D"], w["num_layers"], w["H"], w["P"], w["V"], w["K"]))
    nc = bench.nnet_config(w, 1.0); nc["is_training"] = False
    model = AcousticModel(nc, torch.device("cuda:0"), seed=1234)
    nnet_in = os.path.join(tmp, "nnet.0")
    Saver(model).save(None, nnet_in)
    del model
    prior = os.path.join(tmp, "prior.txt")
    with open(prior, "w") as fh:
        fh.write("[ " + " ".join("%d" % c for c in rng.randint(1, 100, size=w["V"])) + " ]\n")
    out = {}
    for bs in batch_sizes:
        n_run = n_utts if bs > 1 else min(n_utts, 256)
        scp_run = scp
        if n_run < n_utts:
            scp_run = os.path.join(tmp, "feats_head.scp")
            open(scp_run, "w").write("".join(open(scp).readlines()[:n_run]))
        ark = os.path.join(tmp, "post_b%d.ark" % bs)
        t0 = time.perf_counter()
        eng = cli.nnet_forward([scp_run, cfg, nnet_in, "ark:" + ark, "--class-prior", prior, "--batch-size", str(bs),
                                "--report-interval", "0", "--io-threads", "8"])
        wall = time.perf_counter() - t0
        out["batch_size_%d" % bs] = {"utts": n_run, "wall_s": wall, "utts_per_s_end_to_end": n_run / wall,
                                    "device_s": eng.timing["device_s"], "utts_per_s_device_only": n_run / max(eng.timing["device_s"], 1e-9),
                                    "padding_overhead": eng.timing["padded_frames"] / max(eng.timing["frames"], 1) - 1.0,
                                    "ark_bytes": os.path.getsize(ark)}
        os.remove(ark)
        torch.cuda.empty_cache()
    shutil.rmtree(tmp, ignore_errors=True)
    return out


if __name__ == "__main__":
    res = {"workload": "C5: WSJ-shape BiLSTM-MoS forward + log-softmax - prior (nnet-forward.py posterior path)", "dtype": "f16 operands, f32 accumulate/state"}
    for B, T, it in ((512, 700, 5), (64, 700, 10), (1, 700, 20)):
        for e2e in (False, True):
            ms, frames = run(B, T, it, e2e)
            res["B%d_%s" % (B, "e2e" if e2e else "device")] = {"ms_per_batch": ms, "frames_per_s": frames / ms * 1e3, "utts_per_s": B / ms * 1e3}
        torch.cuda.empty_cache()
    if "--cli" in sys.argv:
        res["cli_nnet_forward"] = run_cli(4096, 700, (1, 64, 512))
    print(json.dumps(res))
