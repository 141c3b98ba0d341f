"""
Multi-GPU plumbing (one process per GPU, `torch.distributed`): independent symbol streams are sharded one per rank --
there is no cross-GPU dependency inside a symbol (the reference holds one symbol per TradesData) -- and the finished
bar frames are collected on rank 0 with ONE gather per step.  Works on NCCL (device tensors) and gloo (CPU tensors;
used by the world_size-2 CPU tests).
"""
from typing import List, Optional, Sequence


def shard_symbols(symbols: Sequence, rank: int, world: int) -> List:
    """Round-robin assignment of symbols to ranks (symbol k -> rank k % world)."""
    return [s for k, s in enumerate(symbols) if k % world == rank]


def gather_frames(frame, dst: int = 0) -> Optional[List]:
    """Gather variable-length 1-D uint8 tensors (serialised bar frames) to ``dst``.

    ``torch.distributed.gather`` needs equal sizes, so sizes are all-gathered first and frames are padded to the
    maximum (NCCL has no gather-v; the payload is MBs against 900 GB/s links).  Returns the list of exact-size frames
    on ``dst`` and None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    n = frame.numel()
    sizes = torch.zeros(world, dtype=torch.int64, device=frame.device)
    dist.all_gather_into_tensor(sizes, torch.tensor([n], dtype=torch.int64, device=frame.device))
    mx = int(sizes.max().item())
    padded = torch.zeros(mx, dtype=torch.uint8, device=frame.device)
    padded[:n] = frame
    out = [torch.empty(mx, dtype=torch.uint8, device=frame.device) for _ in range(world)] if rank == dst else None
    dist.gather(padded, out, dst=dst)
    if rank != dst:
        return None
    return [out[r][: int(sizes[r].item())] for r in range(world)]


class PipelinedFrameGather:
    """Gather of variable-length frames to ``dst``, one per step, OVERLAPPED with the next step's compute.

    The frame of step k is copied (device to device, on the caller's stream) into one of two fixed-capacity staging
    buffers -- 16-byte header carrying the exact length, then the payload -- and gathered on a separate communication
    stream; the kernels of step k+1 run meanwhile.  A staging buffer is reused only after the gather that read it has
    finished (event).  No size exchange and no host synchronisation per step: every rank sends ``capacity + 16`` bytes.
    ``finish()`` makes the caller's stream wait for the outstanding gathers, so a timer stopped on that stream covers
    them, and returns the exact-size frames of the last step on ``dst``.  CPU tensors (gloo) take a synchronous path."""

    HDR = 16

    def __init__(self, frame, dst: int = 0, slack: float = 1.25):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.dst = torch, dist, dst
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.cuda = frame.device.type == "cuda"
        cap = torch.tensor([int(frame.numel() * slack) + 1024], dtype=torch.int64, device=frame.device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)          # one agreement on the capacity, at construction
        self.capacity = int(cap.item())
        mk = lambda: torch.zeros(self.capacity + self.HDR, dtype=torch.uint8, device=frame.device)   # noqa: E731
        self.buf = [mk(), mk()]
        self.out = [[mk() for _ in range(self.world)] for _ in range(2)] if self.rank == dst else [None, None]
        self.done = [None, None]
        self.k = 0
        self.last = None
        import os
        self.overlap = os.environ.get("FMK_GATHER_OVERLAP", "1") != "0"
        self.comm = torch.cuda.Stream(frame.device) if (self.cuda and self.overlap) else None

    def submit(self, frame):
        torch, dist = self.torch, self.dist
        n = frame.numel()
        if n > self.capacity:
            raise ValueError(f"frame of {n} bytes exceeds the agreed capacity {self.capacity}")
        s = self.k & 1
        if self.cuda:
            main = torch.cuda.current_stream(frame.device)
            if self.done[s] is not None:
                main.wait_event(self.done[s])               # the gather that read this buffer two steps ago is finished
            self.buf[s][:8].view(torch.int64).fill_(n)
            self.buf[s][self.HDR:self.HDR + n].copy_(frame, non_blocking=True)
            if self.comm is None:                           # same-stream variant: no overlap, no size exchange
                dist.gather(self.buf[s], self.out[s], dst=self.dst)
            else:
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(self.comm):
                    self.comm.wait_event(ready)
                    dist.gather(self.buf[s], self.out[s], dst=self.dst)
                    ev = torch.cuda.Event()
                    ev.record(self.comm)
                self.done[s] = ev
        else:
            self.buf[s][:8].view(torch.int64).fill_(n)
            self.buf[s][self.HDR:self.HDR + n].copy_(frame)
            dist.gather(self.buf[s], self.out[s], dst=self.dst)
        self.last = s
        self.k += 1

    def finish(self):
        """Wait (stream-ordered) for the outstanding gathers; on ``dst`` return the frames of the last step."""
        torch = self.torch
        if self.cuda:
            main = torch.cuda.current_stream(self.buf[0].device)
            for ev in self.done:
                if ev is not None:
                    main.wait_event(ev)
        if self.rank != self.dst or self.last is None:
            return None
        frames = []
        for r in range(self.world):
            o = self.out[self.last][r]
            n = int(o[:8].view(torch.int64).item())
            frames.append(o[self.HDR:self.HDR + n])
        return frames
