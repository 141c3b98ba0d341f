"""feature/core/volatility.py:139-219 (``ewmst``) on the GPU: affine-map scan instead of the serial recurrence."""
from ... import core


def ewmst(timestamps, y, half_life, sigma_floor=1e-12, ctx=None):
    return core.ewmst_series(timestamps, y, half_life, sigma_floor, ctx=ctx)


def ewms(y, span, ctx=None):
    """feature/core/volatility.py:9-69: pandas-equivalent EW std (adjust=True, bias=False)."""
    return core.ewms_series(y, span, ctx=ctx)


def realized_vol(r, window, is_sample, ctx=None):
    """feature/core/volatility.py:256-286: rolling RMS of returns with NaN skipping."""
    return core.realized_vol_series(r, window, is_sample, ctx=ctx)
