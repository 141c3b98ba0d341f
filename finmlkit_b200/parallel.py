"""
Multi-GPU plumbing: one process per GPU, independent symbol streams sharded one per rank -- there is no cross-GPU
dependency inside a symbol (the reference holds one symbol per TradesData) -- and ONE gather of the finished bar frames
to rank 0 per step.  The collective is libfmk's own (csrc/comm.cu: ``ncclAllGather`` of exact byte counts + grouped
``ncclSend`` / ``ncclRecv``, NCCL loaded with dlopen); nothing here imports torch.

Bootstrap: NCCL needs its 128-byte unique id handed from rank 0 to the other ranks once.  Under ``torchrun`` the
environment gives RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT; all ranks of this tier run on ONE node, so
the id travels through a file in a shared temp directory keyed on the job (MASTER_PORT + TORCHELASTIC_RUN_ID), written
atomically by rank 0 and polled by the others.

The host-side framing (``pack_frame`` / ``unpack_frame`` / ``shard_symbols`` / ``gather_frames_with``) is backend-agnostic
and is what the world_size-2 CPU tests drive through gloo (tests/test_parallel_gloo.py); the GPU path is ``Comm``.
"""
import ctypes as C
import os
import struct
import tempfile
import time
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import core

MAGIC = 0x464D4B46   # "FMKF"


def shard_symbols(symbols: Sequence, rank: int, world: int) -> List:
    """Round-robin assignment of symbols to ranks (symbol k -> rank k % world)."""
    return [s for k, s in enumerate(symbols) if k % world == rank]


# ---- framing: a bar frame = named columns -> one byte string (header + 16-byte aligned payloads) ----------------------
def pack_frame(columns: dict) -> np.ndarray:
    """{name: 1-D array} -> uint8 array: magic, n_cols, then per column (name, dtype, count, offset), then payloads."""
    metas, payload, off = [], [], 0
    for name, a in columns.items():
        a = np.ascontiguousarray(a)
        nb, ds = name.encode(), a.dtype.str.encode()
        metas.append(struct.pack("<H", len(nb)) + nb + struct.pack("<H", len(ds)) + ds + struct.pack("<qq", a.size, off))
        pad = (-a.nbytes) % 16
        payload.append(a.view(np.uint8).reshape(-1))
        if pad:
            payload.append(np.zeros(pad, np.uint8))
        off += a.nbytes + pad
    head = struct.pack("<II", MAGIC, len(metas)) + b"".join(metas)
    head += b"\0" * ((-len(head) - 8) % 16)
    head = struct.pack("<q", len(head) + 8) + head
    return np.concatenate([np.frombuffer(head, np.uint8)] + payload) if payload else np.frombuffer(head, np.uint8).copy()


def unpack_frame(buf) -> dict:
    b = np.ascontiguousarray(buf, dtype=np.uint8)
    raw = b.tobytes()
    (hlen,) = struct.unpack_from("<q", raw, 0)
    magic, ncols = struct.unpack_from("<II", raw, 8)
    if magic != MAGIC:
        raise ValueError("not a finmlkit_b200 bar frame")
    p, out = 16, {}
    for _ in range(ncols):
        (ln,) = struct.unpack_from("<H", raw, p); p += 2
        name = raw[p:p + ln].decode(); p += ln
        (ld,) = struct.unpack_from("<H", raw, p); p += 2
        dt = np.dtype(raw[p:p + ld].decode()); p += ld
        cnt, off = struct.unpack_from("<qq", raw, p); p += 16
        out[name] = np.frombuffer(raw, dtype=dt, count=cnt, offset=hlen + off)
    return out


def gather_frames_with(frame: np.ndarray, rank: int, world: int, dst: int, allgather_i64: Callable, sendrecv: Callable):
    """Gather-v of one uint8 frame per rank to ``dst`` in the collective's own shape: exchange the exact byte counts, then
    point-to-point transfers of exactly that many bytes.  ``allgather_i64(x) -> list of world ints``;
    ``sendrecv(buf_or_None, count, src, dst)`` moves ``count`` bytes from ``src`` to ``dst`` and returns the received array
    on ``dst``.  Returns the list of frames on ``dst`` and None elsewhere."""
    sizes = [int(x) for x in allgather_i64(int(frame.size))]
    out = [None] * world
    for r in range(world):
        if r == dst:
            if rank == dst:
                out[r] = frame
            continue
        got = sendrecv(frame if rank == r else None, sizes[r], r, dst)
        if rank == dst:
            out[r] = got
    return out if rank == dst else None


# ---- the GPU path: libfmk's NCCL communicator ---------------------------------------------------------------------------
def _store_path():
    job = f"{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'none')}"
    return os.path.join(os.environ.get("FMK_STORE_DIR", tempfile.gettempdir()), f"fmk_nccl_id_{os.getuid()}_{job}")


def exchange_unique_id(rank: int, world: int, make_id: Callable[[], bytes], timeout_s: float = 300.0, path: str = None) -> bytes:
    """Rank 0 creates the id and publishes it (atomic rename); the others poll for it.  The last reader removes the file."""
    path = path or _store_path()
    if rank == 0:
        for stale in (path, path + ".acks"):
            try:
                os.remove(stale)
            except FileNotFoundError:
                pass
        uid = make_id()
        tmp = f"{path}.tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(uid)
        os.rename(tmp, path)
        return uid
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as f:
                uid = f.read()
            if len(uid) == 128:
                return uid
        except FileNotFoundError:
            pass
        if time.time() - t0 > timeout_s:
            raise TimeoutError(f"rank {rank}: no NCCL unique id at {path} after {timeout_s} s")
        time.sleep(0.01)


class Comm:
    """libfmk's communicator on a Context's device (one per process).  ``Comm.from_env(ctx)`` reads torchrun's variables."""

    def __init__(self, ctx: core.Context, rank: int, world: int, uid: bytes, max_ctas: int = 4):
        self.ctx, self.rank, self.world = ctx, rank, world
        L = ctx._L
        h = C.c_void_p()
        buf = C.create_string_buffer(uid, 128)
        ctx.check(L.fmk_comm_init(ctx.h, buf, rank, world, int(max_ctas), C.byref(h)))
        self.h = h
        self._L = L

    @classmethod
    def from_env(cls, ctx: core.Context, max_ctas: int = 4):
        rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
        L = ctx._L

        def make_id():
            b = C.create_string_buffer(128)
            if L.fmk_comm_unique_id(b) != 0:
                raise core.FmkError("ncclGetUniqueId failed (libnccl.so.2 missing?)")
            return b.raw
        uid = exchange_unique_id(rank, world, make_id)
        c = cls(ctx, rank, world, uid, max_ctas)
        c.barrier()
        if rank == 0:
            try:
                os.remove(_store_path())
            except OSError:
                pass
        return c

    def barrier(self):
        self.ctx.check(self._L.fmk_comm_barrier(self.h))

    @property
    def payload_path(self) -> str:
        """how frames travel: peer-to-peer pushes into the destination's IPC-mapped buffer (copy engines), or NCCL send/recv"""
        return "p2p-ipc-copy-engine" if self._L.fmk_comm_p2p_active(self.h) else "nccl-send-recv"

    def allreduce(self, values, op="max"):
        a = np.ascontiguousarray(values, dtype=np.float64).reshape(-1).copy()
        self.ctx.check(self._L.fmk_comm_allreduce_f64(self.h, a.ctypes.data_as(C.c_void_p), len(a), {"max": 0, "min": 1, "sum": 2}[op]))
        return a

    def gather_submit(self, segments, dst=0):
        """segments: [(device pointer, bytes)] valid on the ctx stream; returns at once (overlaps the next step)."""
        n = len(segments)
        ptrs = (C.c_void_p * max(n, 1))(*[C.c_void_p(p) for p, _ in segments])
        sizes = (C.c_int64 * max(n, 1))(*[int(b) for _, b in segments])
        self.ctx.check(self._L.fmk_comm_gather_submit(self.h, ptrs, sizes, n, int(dst)))

    def gather_finish(self):
        self.ctx.check(self._L.fmk_comm_gather_finish(self.h))

    def gather_reset(self):
        """finish, release the staging / receive buffers, and let the next submit size the pipeline from its own frame
        (collective: every rank calls it)"""
        self.ctx.check(self._L.fmk_comm_gather_reset(self.h))

    def gathered_bytes(self):
        """exact byte count of every rank's frame in the last finished gather"""
        out = []
        for r in range(self.world):
            p, b = C.c_void_p(), C.c_int64()
            self.ctx.check(self._L.fmk_comm_gather_result(self.h, r, C.byref(p), C.byref(b)))
            out.append(int(b.value))
        return out

    def gathered_frame(self, r) -> Optional[np.ndarray]:
        """host copy of rank r's frame of the last step (destination rank only)"""
        n = self.gathered_bytes()[r]
        out = np.empty(n, np.uint8)
        self.ctx.check(self._L.fmk_comm_gather_download(self.h, r, out.ctypes.data_as(C.c_void_p), n))
        return out

    def destroy(self):
        if self.h:
            self._L.fmk_comm_destroy(self.h)
            self.h = None
