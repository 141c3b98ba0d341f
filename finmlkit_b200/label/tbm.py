"""label/tbm.py:11-158 on the GPU (warp per event)."""
import numpy as np

from .. import core


def triple_barrier(timestamps, close, event_idxs, targets, horizontal_barriers, vertical_barrier, min_close_time_sec,
                   side, min_ret, ctx=None, _dev_trades=None):
    """Same arguments, return tuple ``(labels i8, touch_idxs i64, rets f64, max_rb_ratios f64)`` and ValueError texts as
    the reference.  Skipped events (vertical barrier index <= event index) get ``touch_idx = event_idx`` -- the reference
    leaves that element uninitialised (``np.empty``, tbm.py:72)."""
    if vertical_barrier <= 0:
        raise ValueError("The vertical barrier must be greater than zero.")
    if min_ret < 0:
        raise ValueError("The minimum return must be non-negative.")
    if len(timestamps) != len(close):
        raise ValueError("The lengths of timestamps and close must match.")
    if len(event_idxs) != len(targets):
        raise ValueError("The lengths of event_idxs and targets must match.")
    if len(event_idxs) == 0:
        raise ValueError("The event_idxs array must not be empty.")
    if side is not None and len(event_idxs) != len(side):
        raise ValueError("The length of event_idxs must match the length of side.")
    tr = _dev_trades
    if tr is None:
        n = len(close)
        tr = core.DeviceTrades.upload(timestamps, close, np.zeros(n, np.float64), ctx=ctx)
    return core.triple_barrier_dev(tr, event_idxs, targets, horizontal_barriers, vertical_barrier, min_close_time_sec,
                                   side, min_ret)
