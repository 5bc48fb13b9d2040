"""CTC loss + gradient on the GPU -- host-side mirror of the reference's
`tf.nn.ctc_loss(labels, inputs, sequence_length, ignore_longer_outputs_than_inputs=True)` call
(/root/reference/nnet/graph.py:109-114) including the dense(-1 padded)->sparse label handling of
graph.py:74-104.  Same argument meaning and error behaviour: labels outside [0, V-1) raise
InvalidArgumentError; too-long label sequences are silently zeroed; blank = V-1.
"""
import torch

from . import _lib

_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def ctc_loss_grad(logits, labels, seq_len, check_labels=True, lattice_layout=-1):
    """logits [B,T,V] f32 cuda (batch-major), labels [B,Lmax] int64 cuda (-1 padded),
    seq_len [B] int32 cuda.  Returns (loss[B], grad[B,T,V]) -- one pass, TF semantics.
    lattice_layout: -1 chosen by batch size, 0 / 1 force one / two CTAs per utterance (tests, A/B timing)."""
    L = _lib.lib()
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 3
    logits = logits.contiguous()
    B, T, V = logits.shape
    labels = labels.to(device=logits.device, dtype=torch.int64).contiguous().view(B, -1)
    seq_len = seq_len.to(device=logits.device, dtype=torch.int32).contiguous()
    Lmax = labels.shape[1]
    nbytes = L.lcb_ctc_workspace_bytes(B, T, V, Lmax)
    if nbytes == 0:
        raise _lib.LcbError(-3, "lcb_ctc_workspace_bytes(B=%d,T=%d,V=%d,Lmax=%d)" % (B, T, V, Lmax))
    ws = _workspace(nbytes, logits.device)
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits)
    st = L.lcb_ctc_loss_grad_f32_layout(_lib.ptr(logits), _lib.ptr(labels) if Lmax > 0 else None, Lmax, _lib.ptr(seq_len),
                                        B, T, V, _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(ws), ws.numel(), int(lattice_layout),
                                        _lib.stream_ptr())
    _lib.check(st, "lcb_ctc_loss_grad_f32")
    if check_labels:
        _lib.check(L.lcb_ctc_status(_lib.ptr(ws), _lib.stream_ptr()), "tf.nn.ctc_loss labels")
    return loss, grad


class _CTCLoss(torch.autograd.Function):
    """autograd node: backward multiplies the kernel's gradient by the upstream d(loss_b)
    (TF's _CTCLossGrad, python/ops/ctc_ops.py)."""

    @staticmethod
    def forward(ctx, logits, labels, seq_len):
        loss, grad = ctc_loss_grad(logits, labels, seq_len)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout.view(-1, 1, 1), None, None


def ctc_loss(labels, inputs, sequence_length, ignore_longer_outputs_than_inputs=True, time_major=False):
    """Drop-in for the reference's call: returns loss[B] (differentiable w.r.t. inputs).
    `inputs` is [B,T,V] (time_major=False, what create_logits_blstm returns) or [T,B,V]."""
    if not ignore_longer_outputs_than_inputs:
        raise NotImplementedError("the reference always passes ignore_longer_outputs_than_inputs=True (graph.py:113)")
    if time_major:
        inputs = inputs.transpose(0, 1)
    return _CTCLoss.apply(inputs, labels, sequence_length)
