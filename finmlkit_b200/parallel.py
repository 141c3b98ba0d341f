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
