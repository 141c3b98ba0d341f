"""feature/core/utils.py:12-64 on the GPU."""
from ... import core


def comp_lagged_returns(timestamps, close, return_window_sec, is_log, ctx=None):
    """Lagged (log) returns over a time window; same semantics as the reference's ``comp_lagged_returns``."""
    return core.lagged_returns(timestamps, close, return_window_sec, is_log, ctx=ctx)
