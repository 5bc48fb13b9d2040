"""Validation / inference tails: greedy CTC decoding (graph.py:138-142) and the posterior of
create_graph_for_inference + nnet-forward.py (graph.py:236, nnet-forward.py:87-91)."""
import torch

from . import _lib


def greedy_decode(logits, seq_len):
    """logits [B,T,V] f32 cuda -> list of label lists (argmax, collapse repeats, drop blank V-1)."""
    B, T, V = logits.shape
    out = torch.empty(B, T, dtype=torch.int32, device=logits.device)
    n = torch.empty(B, dtype=torch.int32, device=logits.device)
    seq_len = seq_len.to(device=logits.device, dtype=torch.int32).contiguous()
    _lib.check(_lib.lib().lcb_greedy_decode(_lib.ptr(logits.contiguous()), _lib.ptr(seq_len), _lib.ptr(out), _lib.ptr(n),
                                            B, T, V, _lib.stream_ptr()), "lcb_greedy_decode")
    out_h, n_h = out.cpu(), n.cpu()
    return [out_h[b, :int(n_h[b])].tolist() for b in range(B)]


def softmax_rows(logits, smooth_factor=1.0, apply_log=False, log_prior=None, blank_to_front=False, out=None):
    """softmax(smooth_factor * logits) over the last axis; optional log, log-prior subtraction and blank -> column 0 reorder
    (scripts/decode_ctc_lat.sh:161-163)."""
    x = logits.contiguous()
    V = x.shape[-1]
    rows = x.numel() // V
    if out is None:
        out = torch.empty_like(x)
    assert out.is_contiguous() and out.shape == x.shape and out.data_ptr() != x.data_ptr()
    lp = None
    if log_prior is not None:
        lp = torch.as_tensor(log_prior, dtype=torch.float32).to(x.device).contiguous()
    _lib.check(_lib.lib().lcb_posterior(_lib.ptr(x), _lib.ptr(out), rows, V, float(smooth_factor), 1 if apply_log else 0,
                                        _lib.ptr(lp), 1 if blank_to_front else 0, _lib.stream_ptr()), "lcb_posterior")
    return out
