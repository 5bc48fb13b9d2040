"""Single-node data parallelism, one process per GPU (torch.distributed / NCCL over NVLink-NVSwitch).

The reference is single-device; utterances are independent through BiLSTM, mixture layer and CTC, so the
batch shards across ranks with replicated weights and exactly ONE exchange per step: a SUM all-reduce of
the flat gradient buffer (SUM, not mean: the reference loss is a batch sum, graph.py:116, so the reduced
gradient equals the single-device gradient on the concatenated batch).  The flat buffer is ordered by
gradient completion (output layer first, layer 0 last): each bucket is all-reduced on a side stream as
soon as its layer's wgrad GEMMs are enqueued, overlapping the serial BPTT of the layers below."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
    else:
        device = torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world)
    return rank, world, device


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def shard_utterances(num_utts, rank, world):
    """Length-sorted utterances are dealt round-robin so every rank sees the same length profile."""
    return list(range(rank, num_utts, world))


def bucket_bounds(params, names):
    """[lo, hi) slice of the flat buffers covering the (contiguous) variables `names`."""
    specs = [params.specs[n] for n in names]
    lo = min(s.offset for s in specs)
    hi = max(s.offset + s.numel for s in specs)
    return lo, hi


class GradientAllReducer:
    """Bucketed, overlapped all-reduce of `params.gflat`.  Works on any backend (gloo on CPU in tests)."""

    def __init__(self, params, group=None):
        self.params = params
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = params.gflat.is_cuda
        self.comm = torch.cuda.Stream() if (self.cuda and self.world > 1) else None
        self.pending = []
        self.reduced = []            # [(lo, hi)] of this step, for tests

    def broadcast_weights(self, src=0):
        if self.world > 1:
            dist.broadcast(self.params.flat, src=src, group=self.group)

    def begin_step(self):
        self.pending, self.reduced = [], []

    def bucket_ready(self, names):
        if self.world == 1:
            return
        lo, hi = bucket_bounds(self.params, names)
        self.reduced.append((lo, hi))
        buf = self.params.gflat[lo:hi]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                self.pending.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.pending.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Make the compute stream wait for every bucket (global-norm clipping needs the reduced gradient)."""
        for w in self.pending:
            w.wait()
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)
        self.pending = []

    def all_reduce_scalars(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t
