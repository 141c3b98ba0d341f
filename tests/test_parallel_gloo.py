"""CPU, world_size 2: the N>1 host logic of finmlkit_b200.parallel -- symbol sharding, frame packing, the gather-v protocol
(exact byte counts first, then point-to-point transfers of exactly that many bytes) and the unique-id bootstrap through the
job-keyed file.  gloo stands in for NCCL here (test infrastructure: the product's collective is libfmk's own, csrc/comm.cu,
covered on the GPU by tests/test_gpu_comm.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from finmlkit_b200.parallel import exchange_unique_id, gather_frames_with, pack_frame, shard_symbols, unpack_frame


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _allgather_i64(x):
    world = dist.get_world_size()
    out = torch.zeros(world, dtype=torch.int64)
    dist.all_gather_into_tensor(out, torch.tensor([x], dtype=torch.int64))
    return out.tolist()


def _sendrecv(buf, count, src, dst):
    rank = dist.get_rank()
    if rank == src:
        dist.send(torch.from_numpy(np.ascontiguousarray(buf)), dst=dst)
        return None
    if rank == dst:
        t = torch.empty(count, dtype=torch.uint8)
        dist.recv(t, src=src)
        return t.numpy()
    return None


def _frame_of(rank, step):
    """a ragged bar frame whose sizes and contents depend on rank and step (footprint-like CSR included)"""
    nb = 100 + 37 * rank + 11 * step
    off = np.cumsum(np.r_[0, (np.arange(nb) % 7) + 1]).astype(np.int64)
    return {"close_idx": np.arange(nb, dtype=np.int64) * (rank + 2), "open": np.full(nb, 10.5 + rank + step),
            "volume": np.full(nb, 0.25 * (rank + 1), np.float32), "fp_level_offsets": off,
            "fp_buy_vol": np.arange(off[-1], dtype=np.float32) + step, "fp_buy_imb": (np.arange(off[-1]) % 3 == rank)}


def _worker(rank, world, port, store, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = exchange_unique_id(rank, world, lambda: bytes(range(128)), timeout_s=60, path=store)
        q.put(("uid", rank, uid == bytes(range(128))))
        mine = shard_symbols(["BTC", "ETH", "SOL", "XRP", "ADA"], rank, world)
        q.put(("shard", rank, mine))
        ok = True
        for step in range(3):
            frame = pack_frame(_frame_of(rank, step))
            got = gather_frames_with(frame, rank, world, 0, _allgather_i64, _sendrecv)
            if rank == 0:
                ok = ok and len(got) == world
                for r in range(world):
                    cols, exp = unpack_frame(got[r]), _frame_of(r, step)
                    ok = ok and list(cols) == list(exp)
                    for k in exp:
                        ok = ok and cols[k].dtype == np.asarray(exp[k]).dtype and np.array_equal(cols[k], exp[k])
            else:
                ok = ok and got is None
        q.put(("gather", rank, ok))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_pack_unpack_round_trip():
    cols = _frame_of(1, 2)
    cols["empty"] = np.zeros(0, np.float64)
    back = unpack_frame(pack_frame(cols))
    assert list(back) == list(cols)
    for k in cols:
        assert back[k].dtype == np.asarray(cols[k]).dtype and np.array_equal(back[k], cols[k])
    assert len(unpack_frame(pack_frame({}))) == 0


def test_world_size_2_gather_and_sharding(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    store = str(tmp_path / "uid")
    procs = [ctx.Process(target=_worker, args=(r, world, port, store, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(3 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = {g[1]: g[2] for g in got if g[0] == "shard"}
    assert shards[0] == ["BTC", "SOL", "ADA"] and shards[1] == ["ETH", "XRP"]
    for r in range(world):
        assert ("uid", r, True) in got and ("gather", r, True) in got


def test_parallel_module_is_torch_free():
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "finmlkit_b200", "parallel.py")).read()
    assert "import torch" not in src
