"""CPU, world_size 2, gloo: the data-parallel gradient exchange.  Each rank fills the flat gradient buffer
with rank-dependent values, announces buckets in backward order through GradientAllReducer, and must end
with the SUM over ranks (the reference loss is a batch sum, graph.py:116) in every bucket, identical weights
after broadcast, and summed logging scalars."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lstm_ctc_b200.blstm import ParamSpec, ParamStore
    from lstm_ctc_b200.dist import GradientAllReducer, shard_utterances
    specs = [ParamSpec("out/W", (6, 4), True)]
    for i in reversed(range(2)):
        specs += [ParamSpec("L%d/Wx" % i, (8, 3), True), ParamSpec("L%d/bias" % i, (8,), False)]
    ps = ParamStore(specs, torch.device("cpu"))
    ps.flat.copy_(torch.arange(ps.total, dtype=torch.float32) * (rank + 1))
    red = GradientAllReducer(ps)
    red.broadcast_weights(src=0)
    w_ok = bool(torch.equal(ps.flat, torch.arange(ps.total, dtype=torch.float32)))
    red.begin_step()
    ps.gflat.fill_(float(rank + 1))
    red.bucket_ready(["out/W"])
    red.bucket_ready(["L1/Wx", "L1/bias"])
    red.bucket_ready(["L0/Wx", "L0/bias"])
    red.finish()
    expect = float(sum(range(1, world + 1)))
    g_ok = all(bool((ps.g(n) == expect).all()) for n in ps.order)
    order_ok = [lo for lo, _ in red.reduced] == sorted(lo for lo, _ in red.reduced) and len(red.reduced) == 3
    sc = red.all_reduce_scalars(torch.tensor([1.5 * (rank + 1), 10.0], dtype=torch.float64))
    s_ok = sc.tolist() == [1.5 * expect, 10.0 * world]
    shards = shard_utterances(7, rank, world)
    q.put((rank, w_ok, g_ok, order_ok, s_ok, shards))
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    for rank, w_ok, g_ok, order_ok, s_ok, shards in res:
        assert w_ok and g_ok and order_ok and s_ok, (rank, w_ok, g_ok, order_ok, s_ok)
    assert res[0][5] == [0, 2, 4, 6] and res[1][5] == [1, 3, 5]
