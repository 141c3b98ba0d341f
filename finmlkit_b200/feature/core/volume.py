"""feature/core/volume.py:572-641 (``comp_flow_acceleration``, ``vpin``) on the GPU."""
from ... import core


def vpin(volume_buy, volume_sell, window, ctx=None):
    return core.vpin_series(volume_buy, volume_sell, window, ctx=ctx)


def comp_flow_acceleration(volumes, window, recent_periods, ctx=None):
    return core.flow_acceleration_series(volumes, window, recent_periods, ctx=ctx)


# ---- rolling volume profile (feature/core/volume.py:13-456 of the reference) -------------------------------------------
def _csr_of(price_levels, buy_volumes, sell_volumes):
    """Ragged per-bar lists (NumbaList / list / object array of arrays) -> (level_offsets, levels, buy, sell)."""
    import numpy as np
    nb = len(price_levels)
    off = np.zeros(nb + 1, np.int64)
    for i in range(nb):
        off[i + 1] = off[i] + len(price_levels[i])
    cat = lambda xs, dt: (np.concatenate([np.asarray(x, dtype=dt) for x in xs]) if nb else np.zeros(0, dt))  # noqa: E731
    return off, cat(price_levels, np.int32), cat(buy_volumes, np.float32), cat(sell_volumes, np.float32)


def volume_profile_rolling(ts, highs, lows, price_levels, buy_volumes, sell_volumes, window_size_sec, n_bins=None,
                           price_tick=None, va_pct=68.34, ctx=None):
    """volume.py:396-456: ``(poc i32, hva i32, lva i32, vp_pct_abv_poc f32)`` aligned to the bars, on the GPU.  The ragged
    lists are flattened to CSR once; decisions are taken on float32 sums formed in the reference's order (bit-exact)."""
    if not (len(ts) == len(highs) == len(lows) == len(price_levels) == len(buy_volumes) == len(sell_volumes) > 0):
        raise AssertionError("Input arrays should have the same length and be non-empty.")
    off, lv, bv, sv = _csr_of(price_levels, buy_volumes, sell_volumes)
    return core.volume_profile_rolling_csr(ts, highs, lows, off, lv, bv, sv, window_size_sec, n_bins, price_tick, va_pct, ctx=ctx)


class VolumePro:
    """volume.py:13-130: same constructor, ``reset_parameters``, ``compute`` and ``compute_range``."""

    def __init__(self, window_size, n_bins: int = 27, va_pct: float = 68.34):
        self.window_size_sec = window_size.total_seconds()
        self.n_bins = n_bins
        self.va_pct = va_pct

    def reset_parameters(self, window_size_sec=None, n_bins=None, va_pct=None):
        self.window_size_sec = window_size_sec if window_size_sec is not None else self.window_size_sec
        self.n_bins = n_bins if n_bins is not None else self.n_bins
        self.va_pct = va_pct if va_pct is not None else self.va_pct

    def compute(self, bars, fp_data):
        import numpy as np
        assert len(bars) == len(fp_data.bar_timestamps), "Bars and footprint data should have the same length."
        poc, hva, lva, pct = volume_profile_rolling(
            fp_data.bar_timestamps, bars.high.values, bars.low.values, fp_data.price_levels, fp_data.buy_volumes,
            fp_data.sell_volumes, window_size_sec=self.window_size_sec, n_bins=self.n_bins, price_tick=fp_data.price_tick,
            va_pct=self.va_pct)
        poc, hva, lva = poc * fp_data.price_tick, hva * fp_data.price_tick, lva * fp_data.price_tick
        poc = np.where(poc == 0, np.nan, poc)
        hva = np.where(hva == 0, np.nan, hva)
        lva = np.where(lva == 0, np.nan, lva)
        return poc, hva, lva, pct

    def compute_range(self, bars, fp_data, start, end):
        import pandas as pd
        assert len(bars) == len(fp_data.bar_timestamps), "Bars and footprint data should have the same length."
        assert type(start) is type(end), "Start and end should be of the same type."
        if isinstance(start, int):
            end = pd.to_datetime(end)
        start = pd.to_datetime(start)
        adjusted_start = start - pd.Timedelta(seconds=self.window_size_sec)
        fp_sub = fp_data[adjusted_start:end]
        bars_sub = bars.loc[pd.to_datetime(fp_sub.bar_timestamps, unit='ns')]
        poc, hva, lva, pct = self.compute(bars_sub, fp_sub)
        return fp_sub.bar_timestamps, poc, hva, lva, pct
