"""
Bar indexers with the reference's array signatures (finmlkit/bar/logic.py), computed on the GPU.

Each function uploads the columns it needs, runs the device indexer and returns NumPy arrays shaped like the
reference's outputs (the reference returns ``numba.typed.List`` for the threshold indexers; callers convert with
``np.array(..., dtype=np.int64)`` -- bar/kit.py:66,100,134 -- which works on the arrays returned here as well).
"""
import numpy as np

from .. import core


def _trades(ctx, ts=None, price=None, amount=None, n=None):
    n = n if n is not None else len(ts if ts is not None else (price if price is not None else amount))
    z64 = np.zeros(n, np.float64)
    return core.DeviceTrades.upload(ts if ts is not None else np.zeros(n, np.int64),
                                    price if price is not None else z64, amount if amount is not None else z64, ctx=ctx)


def _time_bar_indexer(timestamps, interval_seconds, ctx=None):
    """logic.py:12-51 -> (bar_clock int64[B+1], bar_close_indices int64[B+1])."""
    tr = _trades(ctx, ts=timestamps)
    return core.time_bar_index(tr, interval_seconds).download()


def _tick_bar_indexer(timestamps, threshold, ctx=None):
    """logic.py:54-84."""
    tr = _trades(ctx, ts=timestamps)
    return core.tick_bar_index(tr, threshold).download()[1]


def _volume_bar_indexer(volumes, threshold, ctx=None):
    """logic.py:87-115."""
    tr = _trades(ctx, amount=volumes)
    return core.volume_bar_index(tr, threshold).download()[1]


def _dollar_bar_indexer(prices, volumes, threshold, ctx=None):
    """logic.py:118-149."""
    tr = _trades(ctx, price=prices, amount=volumes)
    return core.dollar_bar_index(tr, threshold).download()[1]


def _cusum_bar_indexer(timestamps, prices, sigma, sigma_floor, sigma_mult, ctx=None):
    """logic.py:152-221.  Like the reference, NaNs of a float64 ``sigma`` array are forward-filled IN PLACE."""
    if not (len(prices) == len(sigma) == len(timestamps)):
        raise ValueError("Prices, timestamps, and sigma arrays must have the same length.")
    tr = _trades(ctx, ts=timestamps, price=prices)
    sg = core.DeviceBuf.upload(tr.ctx, np.ascontiguousarray(sigma, dtype=np.float64))
    idx = core.cusum_bar_index(tr, sg, sigma_floor, sigma_mult).download()[1]
    if isinstance(sigma, np.ndarray) and sigma.dtype == np.float64 and sigma.flags.writeable:
        sigma[...] = sg.download(np.float64, len(sigma))
    return idx


def _imbalance_bar_indexer(timestamps, prices, volumes, threshold):
    """logic.py:224-241: not implemented in the reference either."""
    raise NotImplementedError("Imbalance bar indexer is not implemented yet.")


def _run_bar_indexer(timestamps, prices, volumes, threshold):
    """logic.py:244-261: not implemented in the reference either."""
    raise NotImplementedError("Run bar indexer is not implemented yet.")
