"""Where the activation error of the CUDA path sits, layer by layer, against the fp64 oracle at C3 model dimensions:
max and rms error of every BiLSTM layer's output h (fp16 on the device), of the output-layer pre-activations and of the
logits.  python tools/parity_diag.py [B T]"""
import sys
import torch
sys.path.insert(0, ".")
import oracle  # noqa: E402
from oracle import model as om  # noqa: E402
from lstm_ctc_b200.model import AcousticModel  # noqa: E402

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (33, 64)
dims = dict(input_dim=120, num_layers=5, num_neurons=512, num_projects=512, num_targets=72, use_peepholes=True, num_experts=8)
cfg = oracle.OracleConfig(**dims)
params = oracle.init_params(cfg, seed=101, bias_scale=0.1)
g = torch.Generator().manual_seed(202)
x = torch.randn(B, T, 120, generator=g, dtype=torch.float64)
lens = torch.randint(int(0.6 * T), T + 1, (B,), generator=g).to(torch.int32)
lens[0] = T
for b in range(B):
    x[b, lens[b]:] = 0
dev = torch.device("cuda:0")
nc = {"nnet_type": "blstm", "input_dim": 120, "left_context": 0, "right_context": 0, "num_layers": 5, "num_neurons": 512,
      "num_projects": 512, "num_targets": 72, "use_peepholes": True, "num_experts": 8, "moe_temp": 10.0, "dropout_rate": 1.0}
m = AcousticModel(nc, dev, init=False)
m.from_tf_dict(params)
logits = m.forward_logits(x.float().to(dev), lens.to(dev), training=True).cpu().double()
ws = m.enc._workspace(T, B, True)
live = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1))
# oracle, layer by layer (bilstm.py:170-203)
finput, binput = x, om.reverse_sequence(x, lens)
for i in range(cfg.num_layers):
    fo, _ = om.dynamic_rnn(finput, lens, om._cell_params(params, cfg, i, "fd", "frnn"), cfg.forget_bias)
    bo, _ = om.dynamic_rnn(binput, lens, om._cell_params(params, cfg, i, "bd", "brnn"), cfg.forget_bias)
    cat = torch.cat([fo, om.reverse_sequence(bo, lens)], 2)
    finput, binput = cat, om.reverse_sequence(cat, lens)
    ours = ws["Hout"][i].float().cpu().double().view(T, B, -1).permute(1, 0, 2)
    e = (ours - cat)[live]
    print("layer %d  h: max|ref| %.3f rms|ref| %.4f   max err %.2e  rms err %.2e   (fp16 ulp at 1.0 = 4.9e-4)"
          % (i, cat.abs().max(), cat[live].pow(2).mean().sqrt(), e.abs().max(), e.pow(2).mean().sqrt()))
ref = om.output_layer(params, cfg, finput)
e = (logits - ref)[live]
print("logits: max|ref| %.3f  max err %.3e  rms err %.3e  (max err / scale %.2e)" % (ref.abs().max(), e.abs().max(), e.pow(2).mean().sqrt(), e.abs().max() / ref.abs().max()))
# the same output layer in fp64 fed with OUR encoder output: separates the encoder's error from the output layer's
ours_enc = ws["Hout"][-1].float().cpu().double().view(T, B, -1).permute(1, 0, 2)
ref2 = om.output_layer(params, cfg, ours_enc)
e2 = (logits - ref2)[live]
print("output layer alone (fp64 layer on our encoder output): max err %.3e rms %.3e" % (e2.abs().max(), e2.pow(2).mean().sqrt()))
