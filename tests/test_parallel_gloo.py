"""CPU, world_size 2 over gloo: the N>1 host logic (symbol sharding + variable-length gather of bar frames)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from finmlkit_b200.parallel import PipelinedFrameGather, gather_frames, shard_symbols


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
    try:
        mine = shard_symbols(["BTC", "ETH", "SOL", "XRP", "ADA"], rank, world)
        # a fake serialised bar frame whose length depends on the rank (ragged gather)
        frame = torch.from_numpy(np.full(1000 + 37 * rank, 10 + rank, np.uint8))
        frames = gather_frames(frame, dst=0)
        if rank == 0:
            ok = len(frames) == world and all(f.numel() == 1000 + 37 * r and bool((f == 10 + r).all()) for r, f in enumerate(frames))
            q.put(("gather", ok))
        else:
            assert frames is None
        # pipelined gatherer (synchronous path on CPU tensors): three steps with frames of varying length and content
        g = PipelinedFrameGather(frame, dst=0)
        ok = True
        for step in range(3):
            f = torch.from_numpy(np.full(900 + 50 * step + 37 * rank, 20 + step + rank, np.uint8))
            g.submit(f)
            fr = g.finish()
            if rank == 0:
                ok = ok and len(fr) == world and all(x.numel() == 900 + 50 * step + 37 * r and bool((x == 20 + step + r).all())
                                                       for r, x in enumerate(fr))
            else:
                ok = ok and fr is None
        try:
            g.submit(torch.zeros(g.capacity + 1, dtype=torch.uint8))
            ok = False
        except ValueError:
            pass
        q.put(("pipelined", rank, ok))
        q.put(("shard", rank, mine))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gather_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2 * world + 1)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = {g[1]: g[2] for g in got if g[0] == "shard"}
    assert shards[0] == ["BTC", "SOL", "ADA"] and shards[1] == ["ETH", "XRP"]
    assert ("gather", True) in got
    assert ("pipelined", 0, True) in got and ("pipelined", 1, True) in got
