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


def _imbalance_bar_indexer(timestamps, prices, volumes, threshold, sides=None, ctx=None):
    """Tick-imbalance bars.  The reference's function is a stub that raises ``NotImplementedError`` (logic.py:224-241), so the
    semantics here are this package's own and are checked only against its own CPU oracle (**parity unpinned**):
    ``b_t`` = ``sides`` when given, else the tick rule on ``prices`` (the reference's ``comp_trade_side_vector``,
    bar/utils.py:12-46); the list starts with index 0 like logic.py:73-84; the bar closes at the first tick with
    ``|sum of b_t since the previous close| >= threshold`` and the sum restarts from 0 (AFML 2.3.2.1 with a fixed expected
    imbalance; the stub's signature ``(timestamps, prices, volumes, threshold)`` is kept, ``volumes`` is not read)."""
    return _signed_tick_bars(timestamps, prices, threshold, sides, 0, ctx)


def _run_bar_indexer(timestamps, prices, volumes, threshold, sides=None, ctx=None):
    """Tick-run bars (stub in the reference, logic.py:244-261; own semantics, **parity unpinned**): buys and sells are counted
    since the previous close and the bar closes at the first tick with ``max(#buys, #sells) >= threshold`` (AFML 2.3.2.2 with
    a fixed expected run length)."""
    return _signed_tick_bars(timestamps, prices, threshold, sides, 1, ctx)


def _signed_tick_bars(timestamps, prices, threshold, sides, kind, ctx):
    if len(timestamps) != len(prices):
        raise ValueError("Prices and timestamps arrays must have the same length.")
    n = len(prices)
    tr = core.DeviceTrades.upload(timestamps, prices, np.zeros(n, np.float64),
                                  np.asarray(sides).astype(np.int8) if sides is not None else None, ctx=ctx)
    return core.imbalance_bar_index(tr, threshold, use_side=sides is not None, kind=kind).download()[1]
