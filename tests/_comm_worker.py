"""One rank of the two-GPU communicator test (tests/test_gpu_round2.py): builds frames of rank-dependent size, gathers them
to rank 0 through libfmk's NCCL communicator over several pipelined steps, and checks on rank 0 that what arrived is byte
for byte what each rank packed (the other rank leaves its packed frames in a shared temp directory)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from finmlkit_b200 import core                      # noqa: E402
from finmlkit_b200.parallel import Comm             # noqa: E402
from finmlkit_b200.synth import synth_trades        # noqa: E402


def frame_bytes(ctx, fr):
    bar = np.zeros((fr.bar_bytes + 15) // 16 * 16, np.uint8)
    lvl = np.zeros((fr.level_bytes + 15) // 16 * 16, np.uint8)
    ctx.check(ctx._L.fmk_frame_download(ctx.h, fr.h, bar.ctypes.data_as(C.c_void_p), lvl.ctypes.data_as(C.c_void_p) if fr.level_bytes else None))
    return np.concatenate([bar, lvl]) if fr.level_bytes else bar


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    out_dir = os.environ["FMK_COMM_TEST_DIR"]
    ctx = core.Context(local)
    comm = Comm.from_env(ctx, max_ctas=4)
    ts, px, qty, side = synth_trades(200_000 + 50_000 * rank, seed=100 + rank)
    tr = core.DeviceTrades.upload(ts, px, qty, side, ctx=ctx)
    thresholds = [1e5, 2.5e5, 6e4, 4e5]
    t = comm.allreduce([float(rank + 1)], "sum")
    assert t[0] == world * (world + 1) / 2
    for step, T in enumerate(thresholds):
        ix = core.dollar_bar_index(tr, T * (1 + rank))
        fr = core.bar_features_device(tr, ix, core.F_ALL, price_tick_size=0.1)
        np.save(os.path.join(out_dir, f"frame_r{rank}_s{step}.npy"), frame_bytes(ctx, fr))
        comm.gather_submit(fr.segments(), dst=0)
        del fr, ix
        if step == 1:                       # also exercise finish() in the middle of the pipeline
            comm.gather_finish()
        if step == 2 and os.environ.get("FMK_COMM_TEST_RESET"):      # ... and the collective reset (buffers released, remapped)
            comm.gather_reset()
    comm.gather_finish()
    ctx.sync()
    comm.barrier()
    if rank == 0:
        last = len(thresholds) - 1
        sizes = comm.gathered_bytes()
        for r in range(world):
            exp = np.load(os.path.join(out_dir, f"frame_r{r}_s{last}.npy"))
            got = comm.gathered_frame(r)
            assert sizes[r] == len(exp), (r, sizes, len(exp))
            assert np.array_equal(got, exp), f"rank {r}: gathered bytes differ from the frame that rank packed"
        print("GATHER_OK", comm.payload_path, sizes, flush=True)
    comm.barrier()
    comm.destroy()
    time.sleep(0.1)


if __name__ == "__main__":
    main()
