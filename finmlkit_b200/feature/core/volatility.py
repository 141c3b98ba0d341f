"""feature/core/volatility.py:139-219 (``ewmst``) on the GPU: affine-map scan instead of the serial recurrence."""
from ... import core


def ewmst(timestamps, y, half_life, sigma_floor=1e-12, ctx=None):
    return core.ewmst_series(timestamps, y, half_life, sigma_floor, ctx=ctx)
