"""world_size-2 gloo test of the N>1 plumbing (weights broadcast once, chunks sharded, results gathered)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from thewhisper_b200.parallel import broadcast_weights, gather_results, shard_range, stream_owner

    w = None
    if rank == 0:
        g = torch.Generator().manual_seed(0)
        w = {"enc.0.wqkv": torch.randn(12, 4, generator=g).to(torch.bfloat16), "enc.0.bqkv": torch.randn(12, generator=g)}
    import torch.distributed.distributed_c10d as c10d

    calls = {"n": 0}
    orig = c10d.broadcast

    def counting(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)

    dist.broadcast = counting
    w = broadcast_weights(w, torch.device("cpu"))
    dist.broadcast = orig
    assert calls["n"] == 1, calls  # ONE tensor collective for all weights
    assert w["enc.0.wqkv"].dtype == torch.bfloat16 and tuple(w["enc.0.wqkv"].shape) == (12, 4)
    mine = list(shard_range(7, rank, world))
    local = [(i, float(w["enc.0.bqkv"].sum()) + i) for i in mine]
    allr = gather_results(local)
    if rank == 0:
        q.put((sorted(x for r in allr for x in r), [stream_owner(s, world) for s in range(5)], {k: v.float().sum().item() for k, v in w.items()}))
    else:
        q.put({k: v.float().sum().item() for k, v in w.items()})
    dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    root = next(o for o in outs if isinstance(o, tuple))
    other = next(o for o in outs if isinstance(o, dict))
    results, owners, sums = root
    assert [i for i, _ in results] == list(range(7))  # every chunk exactly once
    assert owners == [0, 1, 0, 1, 0]
    assert other == sums  # identical weights on both ranks
    base = sums["enc.0.bqkv"]
    assert all(abs(v - (base + i)) < 1e-5 for i, v in results)
