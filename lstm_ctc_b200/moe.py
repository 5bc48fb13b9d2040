"""`create_moe` with the reference's signature (/root/reference/nnet/moe.py:29-30):
    create_moe(lstm_output [N,2P], output_dim, num_targets, num_experts, moe_temperature, dropout_rate) -> y [N,V]
Runs the fused sm_100a output kernel.  The four variables the reference creates inside (W_prior, b_prior,
W, b; moe.py:34-58) are passed explicitly in the reference's layout."""
import torch

from . import _lib


def create_moe(lstm_output, output_dim, num_targets, num_experts, moe_temperature, dropout_rate,
               W_prior=None, b_prior=None, W=None, b=None, seed=777):
    keep = 1.0 if dropout_rate is None else float(dropout_rate)
    K, V, D = int(num_experts), int(num_targets), int(output_dim)
    dev = lstm_output.device
    N = lstm_output.shape[0]
    Wall = torch.empty(K * V + K, D, dtype=torch.float16, device=dev)
    Wall[:K * V] = W.t().reshape(K, V, D).permute(1, 0, 2).reshape(K * V, D).to(torch.float16)
    Wall[K * V:] = W_prior.t().to(torch.float16)
    ball = torch.cat([b.view(K, V).t().reshape(K * V), b_prior]).float().contiguous()
    X = lstm_output.to(torch.float16).contiguous()
    y = torch.empty(1, N, V, dtype=torch.float32, device=dev)       # T = N "frames" of a single utterance
    _lib.check(_lib.lib().lcb_output_fwd(_lib.ptr(X), X.stride(0), _lib.ptr(Wall), _lib.ptr(ball), _lib.ptr(y),
                                         N, 1, D, V, K, float(moe_temperature), keep, int(seed), _lib.stream_ptr()), "lcb_output_fwd")
    return y[0]
