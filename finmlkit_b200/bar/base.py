"""
Bar builder base class and the array-level "core functions" of finmlkit/bar/base.py, computed on the GPU.

``BarBuilderBase`` keeps the reference's contract (base.py:24-300): subclasses implement ``_comp_bar_close()`` returning
``(close_ts int64[B+1], close_idx int64[B+1])``; ``build_ohlcv / build_directional_features / build_trade_size_features /
build_footprints`` return frames / ``FootprintData`` with the reference's columns, dtypes, order and index.  The trade
columns are uploaded ONCE per builder into device SoA buffers and every ``build_*`` call runs against them.
"""
import io
from abc import ABC, abstractmethod
from typing import Optional, Tuple

import numpy as np
import pandas as pd

from .. import core
from .data_model import FootprintData
from .utils import comp_price_tick_size


def _device_trades(prices=None, amounts=None, sides=None, n=None, ctx=None):
    if n is None:
        n = len(prices if prices is not None else amounts)
    z = np.zeros(n, np.float64)
    return core.DeviceTrades.upload(np.zeros(n, np.int64), prices if prices is not None else z,
                                    amounts if amounts is not None else z, sides, ctx=ctx)


# ----------------------------------------------------------------------------------------------------------------------
# array-level functions (reference: base.py:303-850) -- same arguments, same return tuples, same ValueErrors
# ----------------------------------------------------------------------------------------------------------------------
def comp_bar_ohlcv(prices, volumes, bar_close_indices, ctx=None):
    """base.py:306-407 -> (open, high, low, close, volume f32, vwap, trades i64, median_trade_size)."""
    if len(prices) != len(volumes):
        raise ValueError("Prices and volumes arrays must have the same length.")
    if len(bar_close_indices) < 2:
        raise ValueError("Bar close indices must contain at least two elements.")
    tr = _device_trades(prices, volumes, ctx=ctx)
    return core.bar_ohlcv(tr, core.DeviceIndex.from_host(tr, bar_close_indices))


def comp_bar_directional_features(prices, volumes, bar_close_indices, trade_sides, ctx=None):
    """base.py:409-546 -> the reference's 14-tuple."""
    tr = _device_trades(prices, volumes, np.asarray(trade_sides).astype(np.int8), ctx=ctx)
    return core.bar_directional(tr, core.DeviceIndex.from_host(tr, bar_close_indices))


def comp_bar_trade_size_features(amounts, theta, bar_close_indices, theta_mult, ctx=None):
    """base.py:549-612 -> (mean_size_rel, size_95_rel, pct_block, size_gini) float32."""
    if len(theta) != len(bar_close_indices) - 1:
        raise ValueError("Theta should match the the number of bars (len(bar_close_indices) - 1).")
    tr = _device_trades(None, amounts, ctx=ctx)
    # the reference takes each bar as the SLICE amounts[start:end + 1] (base.py:585), and Python slices clamp: a close index
    # of len(amounts) -- the reference's own tests use one -- means "through the last trade"
    idx = np.minimum(np.asarray(bar_close_indices, dtype=np.int64), len(amounts) - 1)
    return core.bar_trade_size(tr, core.DeviceIndex.from_host(tr, idx), theta, theta_mult)


def comp_bar_footprints(prices, amounts, bar_close_indices, trade_sides, price_tick_size, bar_lows, bar_highs,
                        imbalance_factor, ctx=None):
    """base.py:615-752 -> 7 ragged lists (one array per bar) + 6 per-bar arrays, in the reference's order."""
    if len(bar_close_indices) < 2:       # no bars: the reference's loop does not run and every output is empty (base.py:676)
        e = lambda dt: np.zeros(0, dt)   # noqa: E731
        return ([], [], [], [], [], [], [], e(np.uint16), e(np.uint16), e(np.int32), e(np.int16), e(np.float64), e(np.float64))
    tr = _device_trades(prices, amounts, np.asarray(trade_sides).astype(np.int8), ctx=ctx)
    csr = core.bar_footprints_csr(tr, core.DeviceIndex.from_host(tr, bar_close_indices), price_tick_size, bar_lows,
                                  bar_highs, imbalance_factor)
    off = csr[0]
    ragged = [[x[off[i]:off[i + 1]] for i in range(len(off) - 1)] for x in csr[1:8]]
    return (*ragged, *csr[8:])


def comp_footprint_features(price_levels, buy_volumes, sell_volumes, imbalance_multiplier, ctx=None):
    """base.py:755-850 for ONE bar's level table -> (buy_imb, sell_imb, imb_max_run_signed, cot, vp_skew, vp_gini).
    Runs the same device kernel as ``comp_bar_footprints`` on a synthetic one-bar stream (one tick per non-zero cell)."""
    lv = np.asarray(price_levels)
    L = len(lv)
    if L == 0:
        return np.zeros(0, np.bool_), np.zeros(0, np.bool_), 0, 0, 0.0, 0.0
    b, s = np.asarray(buy_volumes, np.float32), np.asarray(sell_volumes, np.float32)
    px = np.concatenate([[float(lv[0])], lv.astype(np.float64), lv.astype(np.float64)])
    am = np.concatenate([[0.0], b.astype(np.float64), s.astype(np.float64)])
    sd = np.concatenate([[0], np.ones(L), -np.ones(L)]).astype(np.int8)
    tr = _device_trades(px, am, sd, ctx=ctx)
    ix = core.DeviceIndex.from_host(tr, np.array([0, 2 * L], np.int64))
    csr = core.bar_footprints_csr(tr, ix, 1.0, np.array([float(lv[0])]), np.array([float(lv[-1])]), imbalance_multiplier)
    return csr[6].astype(np.bool_), csr[7].astype(np.bool_), int(csr[11][0]), int(csr[10][0]), float(csr[12][0]), float(csr[13][0])


# ----------------------------------------------------------------------------------------------------------------------
class BarBuilderBase(ABC):
    """Template for bar builders (reference: base.py:24-300)."""

    def __init__(self, trades, ctx=None):
        # ``trades``: a TradesData (this package's or the reference's: only ``.data`` is read, base.py:71) or a
        # ``bar.io.StoreTrades`` whose columns already sit on the device (its pandas frame is then never built here)
        self._trades = trades
        self._store = trades if hasattr(trades, "device_trades") else None
        self._frame = None if self._store is not None else trades.data
        self._ctx = ctx or core.default_context()
        self._close_ts: Optional[np.ndarray] = None
        self._close_indices: Optional[np.ndarray] = None
        self._highs: Optional[np.ndarray] = None
        self._lows: Optional[np.ndarray] = None
        self._dev_trades = None
        self._dev_index = None

    def __str__(self) -> str:
        members = "\n".join(f"{k}: {v}" for k, v in self.__dict__.items())
        buf = io.StringIO()
        try:
            self.trades_df.info(buf=buf)
            info = buf.getvalue()
        except Exception:
            info = "<unavailable>"
        return f"Class: {self.__class__.__name__} with members:\n{members}\nRaw trades data:\n{info}"

    @property
    def trades_df(self) -> pd.DataFrame:
        if self._frame is None:
            self._frame = self._trades.data
        return self._frame

    def _has_side(self) -> bool:
        return self._store.has_side if self._store is not None else 'side' in self.trades_df.columns

    # -- device residency -------------------------------------------------------------------------------------------
    _needs_device_ts = False   # time / CUSUM kits override: their indexers read the timestamps on the device

    def _host_ts(self):
        return self.trades_df['timestamp'].astype(np.int64).values

    def _device(self, need_side: bool = False) -> core.DeviceTrades:
        """The device SoA copy of the trade columns, uploaded ONCE per trades frame and shared with every other builder,
        transform and label call on the same frame (core.device_trades_for).  Price and amount go up first; timestamps only
        for the kits whose indexer reads them on the device, the side column when a directional / footprint build asks."""
        if self._store is not None:
            return self._store.device_trades(need_ts=self._needs_device_ts, need_side=need_side, ctx=self._ctx)
        if self._dev_trades is None or (need_side and not getattr(self._dev_trades, "has_side", True)):
            self._dev_trades = core.device_trades_for(self.trades_df, need_ts=self._needs_device_ts, need_side=need_side,
                                                      ctx=self._ctx)
        return self._dev_trades

    def _download_index(self):
        on_device = self._needs_device_ts or self._store is not None or getattr(self._device(), "has_ts", False)
        return self._dev_index.download(host_ts=None if on_device else self._host_ts())

    @abstractmethod
    def _comp_bar_close(self) -> Tuple[np.ndarray, np.ndarray]:
        """Return (close timestamps, close indices); implementations may also set ``self._dev_index``."""

    def _set_bar_close(self):
        if self._close_ts is None and self._close_indices is None:
            self._close_ts, self._close_indices = self._comp_bar_close()

    def _index(self) -> core.DeviceIndex:
        self._set_bar_close()
        if self._dev_index is None:   # subclass computed the indices on the host (MockBarBuilder-style)
            self._dev_index = core.DeviceIndex.from_host(self._device(), self._close_indices)
        return self._dev_index

    @property
    def bar_close_indices(self):
        if self._close_indices is None:
            self._set_bar_close()
        return self._close_indices[1:]

    @property
    def bar_close_timestamps(self):
        if self._close_ts is None:
            self._set_bar_close()
        return self._close_ts[1:]

    # -- builders -----------------------------------------------------------------------------------------------------
    def build_ohlcv(self) -> pd.DataFrame:
        """base.py:132-169."""
        ix = self._index()
        t = core.bar_ohlcv(self._device(), ix)
        self._highs, self._lows = t[1], t[2]
        df = pd.DataFrame({'timestamp': self.bar_close_timestamps, 'open': t[0], 'high': t[1], 'low': t[2], 'close': t[3],
                           'volume': t[4], 'trades': t[6], 'median_trade_size': t[7], 'vwap': t[5]})
        df['timestamp'] = pd.to_datetime(df['timestamp'], unit='ns')
        df.set_index('timestamp', inplace=True)
        if hasattr(self, 'interval'):
            df.index.freq = pd.Timedelta(seconds=self.interval)
        return df

    def build_directional_features(self) -> pd.DataFrame:
        """base.py:171-212."""
        ix = self._index()
        if not self._has_side():
            raise KeyError('side')
        d = core.bar_directional(self._device(need_side=True), ix)
        names = ['ticks_buy', 'ticks_sell', 'volume_buy', 'volume_sell', 'dollars_buy', 'dollars_sell', 'mean_spread',
                 'max_spread', 'cum_ticks_min', 'cum_ticks_max', 'cum_volume_min', 'cum_volume_max', 'cum_dollars_min',
                 'cum_dollars_max']
        df = pd.DataFrame({'timestamp': self.bar_close_timestamps, **{nm: d[k] for k, nm in enumerate(names)}})
        df['timestamp'] = pd.to_datetime(df['timestamp'], unit='ns')
        df.set_index('timestamp', inplace=True)
        return df

    def build_trade_size_features(self, theta, theta_mult: float = 5.0) -> pd.DataFrame:
        """base.py:214-245."""
        ix = self._index()
        if len(theta) != len(self._close_indices) - 1:
            raise ValueError("Theta should match the the number of bars (len(bar_close_indices) - 1).")
        t = core.bar_trade_size(self._device(), ix, theta, theta_mult)
        df = pd.DataFrame({'timestamp': self.bar_close_timestamps, 'mean_size_rel': t[0], 'size_95_rel': t[1],
                           'pct_block': t[2], 'size_gini': t[3]})
        df['timestamp'] = pd.to_datetime(df['timestamp'], unit='ns')
        df.set_index('timestamp', inplace=True)
        return df

    def build_footprints(self, price_tick_size=None, imbalance_factor=3.0) -> FootprintData:
        """base.py:247-300."""
        ix = self._index()
        if self._highs is None or self._lows is None:
            self.build_ohlcv()
        if price_tick_size is None:
            price_tick_size = comp_price_tick_size(self._store.column("price")[:10000] if self._store is not None
                                                   else self.trades_df['price'].values)
        csr = core.bar_footprints_csr(self._device(need_side=True), ix, price_tick_size, self._lows, self._highs, imbalance_factor)
        fp = FootprintData.from_csr(self.bar_close_timestamps, price_tick_size, csr)
        fp.cast_to_numba_list()
        return fp
