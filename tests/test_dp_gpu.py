"""On-hardware data-parallel equivalence: the SUM all-reduced gradient of a batch sharded over 2 ranks equals the 1-rank
gradient of the concatenated batch -- summed loss (/root/reference/nnet/graph.py:116), gradient, global norm (graph.py:190)
and the weights after one clipped Adam step.

Runs as two processes that BOTH use cuda:0 (the driver's GPU tests see one device): every FLOP is this library's kernels, the
exchange goes through the same GradientAllReducer / bucket order as the multi-GPU step, over gloo (NCCL refuses two ranks on
one device; the N-GPU NCCL path is checked by bench.py's `dp_check` record at N = 2/4/8).  keep_prob = 1: dropout masks are
indexed by the element's position in the LOCAL batch, so sharded and unsharded runs draw different masks by design.

Tolerances: gradient 1e-4 normwise (identical per-utterance activations; only the fp32 summation order of the weight gradients
differs), loss / global norm 1e-5 relative, weights after the update 2e-4 absolute = 0.2 lr (Adam's first step is lr * g / |g|: a
gradient entry below the summation noise may flip sign, damped by eps)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import oracle
from lstm_ctc_b200 import dist as lcb_dist
from lstm_ctc_b200.model import AcousticModel
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
nc = {"nnet_type": "blstm", "input_dim": 40, "left_context": 0, "right_context": 0, "num_layers": 3, "num_neurons": 512,
      "num_projects": 512, "num_targets": 72, "use_peepholes": True, "num_experts": 8, "moe_temp": 10.0, "dropout_rate": 1.0}
B, T = 34, 48
g = torch.Generator().manual_seed(3)
x = torch.randn(B, T, 40, generator=g)
lens = torch.randint(T // 2, T + 1, (B,), generator=g).to(torch.int32); lens[0] = T
for b in range(B):
    x[b, lens[b]:] = 0
y = torch.randint(0, 71, (B, 5), generator=g)
m = AcousticModel(nc, dev, seed=77)
red = lcb_dist.GradientAllReducer(m.params)
red.broadcast_weights()
w0 = m.params.flat.clone()
idx = lcb_dist.shard_utterances(B, rank, world)
red.begin_step()
loss_s, _ = m.loss_and_grad(x[idx].to(dev), lens[idx].to(dev), y[idx].to(dev), bucket_ready=red.bucket_ready)
red.finish()
sc = red.all_reduce_scalars(torch.tensor([float(loss_s)], dtype=torch.float64, device=dev))
torch.cuda.synchronize()
g_dp = m.params.gflat.double().clone()
m.optimizer_step("adam", 1e-3, clip_norm=5.0, l2_decay_weight=1e-5)
n_dp = m.last_grad_norm()
w_dp = m.params.flat.clone()
out = {"rank": rank, "buckets": len(red.reduced)}
if rank == 0:
    m1 = AcousticModel(nc, dev, seed=77)
    m1.params.flat.copy_(w0); m1.mark_stale()
    loss_1, _ = m1.loss_and_grad(x.to(dev), lens.to(dev), y.to(dev))
    g_1 = m1.params.gflat.double().clone()
    m1.optimizer_step("adam", 1e-3, clip_norm=5.0, l2_decay_weight=1e-5)
    out.update(grad_rel=float((g_dp - g_1).norm() / g_1.norm()), loss_dp=float(sc[0]), loss_1=float(loss_1),
               norm_dp=n_dp, norm_1=m1.last_grad_norm(), w_abs=float((w_dp - m1.params.flat).abs().max()),
               w_moved=float((w_dp - w0).abs().max()))
print("DPRESULT " + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gradient_equals_one_rank_gradient(cuda_dev, tmp_path):
    import json
    script = tmp_path / "dp_worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = 29711 + (os.getpid() % 200)
    procs = []
    for r in range(2):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r), WORLD_SIZE="2", LOCAL_RANK="0")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("data-parallel worker timed out")
        outs.append(o)
        assert p.returncode == 0, o[-3000:]
    res = {}
    for o in outs:
        for line in o.splitlines():
            if line.startswith("DPRESULT "):
                d = json.loads(line[9:])
                res[d["rank"]] = d
    r0 = res[0]
    print(r0)
    assert r0["buckets"] == 4                      # output layer + 3 BiLSTM layers, in backward order
    assert r0["grad_rel"] < 1e-4, r0
    assert abs(r0["loss_dp"] - r0["loss_1"]) <= 1e-5 * abs(r0["loss_1"]), r0
    assert abs(r0["norm_dp"] - r0["norm_1"]) <= 1e-5 * r0["norm_1"], r0
    assert r0["w_moved"] > 1e-4 and r0["w_abs"] <= 2e-4, r0
