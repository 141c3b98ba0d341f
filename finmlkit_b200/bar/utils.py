"""Host-side helpers of the bar path (finmlkit/bar/utils.py).  ``comp_price_tick_size`` looks at <= 10 000 prices, so it
stays on the host -- it is NumPy arithmetic, not part of the accelerated stream."""
import math

import numpy as np
import pandas as pd


def comp_price_tick_size(prices) -> float:
    """bar/utils.py:49-81: GCD tick inference on the first 10 000 prices."""
    prices = np.asarray(prices, dtype=np.float64)
    if len(prices) == 0:
        raise ValueError("Empty prices array")
    sample = np.round(prices[:min(10000, len(prices))], decimals=12)
    uniq = np.unique(sample)
    if len(uniq) <= 1:
        return 0.0
    diffs = np.diff(uniq)
    scale = 10.0 ** (-np.floor(np.log10(np.min(diffs[diffs > 0]))))
    int_px = np.round(uniq * scale).astype(np.int64)
    tick_int = 0
    for d in np.diff(int_px):
        d = int(d)
        if d > 0:
            tick_int = d if tick_int == 0 else math.gcd(tick_int, d)
            if tick_int == 1:
                break
    return tick_int / scale


def footprint_to_dataframe(bar_timestamps, price_levels, buy_volumes, sell_volumes, buy_ticks, sell_ticks, buy_imbalance,
                           sell_imbalance, price_tick):
    """Long-format footprint frame, same columns / MultiIndex / ordering as bar/utils.py:129-209 (built vectorised)."""
    n_levels = np.array([len(x) for x in price_levels], dtype=np.int64)
    bar_ids = np.repeat(np.arange(len(n_levels)), n_levels)
    bar_dt = pd.to_datetime(np.asarray(bar_timestamps))

    def cat(xs):
        return np.concatenate([np.asarray(x) for x in xs]) if len(xs) else np.zeros(0)

    data = {
        'price_level': cat(price_levels),
        'sell_ticks': cat(sell_ticks),
        'buy_ticks': cat(buy_ticks),
        'sell_volume': cat(sell_volumes),
        'buy_volume': cat(buy_volumes),
        'sell_imbalance': cat(sell_imbalance),
        'buy_imbalance': cat(buy_imbalance),
    }
    multi_index = pd.MultiIndex.from_arrays([bar_ids, bar_dt[bar_ids]], names=['bar_idx', 'bar_datetime_idx'])
    df = pd.DataFrame(data, index=multi_index)
    df['price_level'] = df['price_level'] * price_tick
    df = df.sort_values(by=['bar_datetime_idx', 'price_level'], ascending=[True, False])
    return df


def comp_trade_side_vector(prices, ctx=None):
    """bar/utils.py:26-46 of the reference (tick rule) on the GPU: int8 sides, ``sides[0] == 0``."""
    from .. import core
    return core.trade_side_vector_dev(prices, ctx=ctx)


def merge_split_trades(timestamps, prices, amounts, is_buyer_maker, ctx=None):
    """bar/utils.py:263-329 of the reference on the GPU.  Inputs must already be ordered by (timestamp, price, side);
    returns ``(timestamps i64, prices f64, amounts f32, sides i8 or empty)``."""
    from .. import core
    return core.merge_split_trades_dev(timestamps, prices, amounts, is_buyer_maker, ctx=ctx)
