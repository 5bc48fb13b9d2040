"""End-to-end (host buffers, Session.run, loss read back every step) C3 step time against the schedule options -- the e2e path
synchronises every step, so host-side enqueue order matters there in a way it does not for the device-resident loop."""
import sys
import json
import types
import torch
sys.path.insert(0, ".")
import bench
import lstm_ctc_b200 as nnet
from lstm_ctc_b200 import blstm, model as model_mod

w = bench.WORKLOADS["c3"]
dev = torch.device("cuda:0")
x_h, lens_h, y_h = bench.synth_batch(w, 777)
frames = float(lens_h.sum())
args = types.SimpleNamespace(steps=10, warmup=4)
orig_enc_init = blstm.BLSTMEncoder.__init__
orig_model_init = model_mod.AcousticModel.__init__
configs = {"default": {}, "fwd_range_launches": {"fwd_flow_control": False}, "no_top_overlap": {"top_overlap": False},
           "bwd_range_launches": {"bwd_progress": False}, "all_off": {"fwd_flow_control": False, "top_overlap": False, "bwd_progress": False},
           "default_again": {}}
for name, cfg in configs.items():
    def enc_init(self, *a, _cfg=cfg, **k):
        orig_enc_init(self, *a, **k)
        for kk, v in _cfg.items():
            if hasattr(self, kk):
                setattr(self, kk, v)

    def model_init(self, *a, _cfg=cfg, **k):
        orig_model_init(self, *a, **k)
        if "top_overlap" in _cfg:
            self.top_overlap = _cfg["top_overlap"]
    blstm.BLSTMEncoder.__init__ = enc_init
    model_mod.AcousticModel.__init__ = model_init
    r = bench.run_e2e(nnet, bench.nnet_config(w, 0.9), w, x_h, lens_h, y_h, args, 1, dev, frames)
    print(json.dumps({"config": name, "settings": cfg, "e2e_ms_per_step": round(r["ms_per_step"], 3)}), flush=True)
    torch.cuda.empty_cache()
