"""
Month-partitioned columnar trade store and its loader straight into device SoA columns.

The reference persists trades as monthly HDF5 tables ``/trades/YYYY-MM`` with ``/meta/YYYY-MM`` records (first / last
timestamp, row count) and reads a time range back by concatenating the months on the host through a process pool
(bar/data_model.py:420-574 ``save_h5`` / ``load_trades_h5``); ``AddTimeBarH5`` then builds 1-second klines per month with
``TimeBarKit`` (bar/io.py:441-514).  PyTables is not part of this image, and the compute path does not want a pandas frame at
all, so the same layout is kept as plain columnar files::

    <root>/<symbol>/trades/YYYY-MM/{timestamp.npy (int64 ns), price.npy (float64), amount.npy (float32 | float64), side.npy (int8)}
    <root>/<symbol>/trades/YYYY-MM/meta.json      first_timestamp, last_timestamp, record_count   (the reference's /meta keys)
    <root>/<symbol>/klines/YYYY-MM/*.npy + meta.json                                             (add_time_bars)

``load_trades_device`` memory-maps the months that intersect ``[start_time, end_time]`` (same discovery rule as
``_keys_for_timerange``, data_model.py:578-593), allocates ONE device handle for the whole range and writes every month's
columns into it at their offset -- the page cache feeds the staged multi-threaded H2D copy directly, no concatenated frame
and no pandas object is built.  ``StoreTrades`` wraps the result so that every bar kit, the sigma transforms and ``TBMLabel``
accept it where they accept a ``TradesData`` (the pandas frame is materialised lazily, only if something asks for it).
"""
import json
import os
from typing import List, Optional

import numpy as np
import pandas as pd

from .. import core

_COLS = (("timestamp", np.int64), ("price", np.float64), ("amount", None), ("side", np.int8))


def _month_key(ts_ns: int) -> str:
    t = pd.Timestamp(int(ts_ns), unit="ns")
    return f"{t.year:04d}-{t.month:02d}"


def save_trades_store(root: str, symbol: str, timestamps, prices, amounts, sides=None, overwrite_month: bool = True) -> List[str]:
    """Write trade columns into the month-partitioned store (one directory per calendar month of the data, like the
    reference's ``/trades/YYYY-MM`` keys).  ``amounts`` keeps its dtype (float32 after a split-trade merge, else float64).
    Returns the month keys written."""
    ts = np.ascontiguousarray(timestamps, dtype=np.int64)
    px = np.ascontiguousarray(prices, dtype=np.float64)
    am = np.ascontiguousarray(amounts)
    if am.dtype not in (np.float32, np.float64):
        am = am.astype(np.float64)
    sd = np.ascontiguousarray(sides, dtype=np.int8) if sides is not None else None
    if not (len(ts) == len(px) == len(am)) or (sd is not None and len(sd) != len(ts)):
        raise ValueError("timestamps, prices, amounts (and sides) must have the same length")
    if len(ts) == 0:
        return []
    if np.any(np.diff(ts) < 0):
        raise ValueError("timestamps must be non-decreasing")
    # month boundaries: first index of every calendar month present in the data
    months = ts.astype("datetime64[ns]").astype("datetime64[M]")
    cut = np.flatnonzero(np.r_[True, months[1:] != months[:-1]])
    bounds = np.r_[cut, len(ts)]
    keys = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        key = _month_key(ts[a])
        d = os.path.join(root, symbol, "trades", key)
        if os.path.isdir(d) and not overwrite_month:
            raise ValueError(f"month {key} already exists in the store")
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, "timestamp.npy"), ts[a:b])
        np.save(os.path.join(d, "price.npy"), px[a:b])
        np.save(os.path.join(d, "amount.npy"), am[a:b])
        if sd is not None:
            np.save(os.path.join(d, "side.npy"), sd[a:b])
        elif os.path.exists(os.path.join(d, "side.npy")):
            os.remove(os.path.join(d, "side.npy"))
        with open(os.path.join(d, "meta.json"), "w") as f:
            json.dump({"first_timestamp": int(ts[a]), "last_timestamp": int(ts[b - 1]), "record_count": int(b - a),
                       "amount_dtype": str(am.dtype), "has_side": sd is not None}, f)
        keys.append(key)
    return keys


def list_months(root: str, symbol: str, start_time=None, end_time=None) -> List[dict]:
    """Months whose ``[first_timestamp, last_timestamp]`` intersects ``[start_time, end_time]`` (data_model.py:578-593)."""
    base = os.path.join(root, symbol, "trades")
    if not os.path.isdir(base):
        raise KeyError(f"no trades for {symbol!r} under {root!r}")
    start = pd.Timestamp(start_time).value if start_time is not None else None
    end = pd.Timestamp(end_time).value if end_time is not None else None
    out = []
    for key in sorted(os.listdir(base)):
        mp = os.path.join(base, key, "meta.json")
        if not os.path.isfile(mp):
            continue
        meta = json.load(open(mp))
        if (end is None or meta["first_timestamp"] <= end) and (start is None or meta["last_timestamp"] >= start):
            meta["key"], meta["dir"] = key, os.path.join(base, key)
            out.append(meta)
    if not out:
        raise KeyError("no monthly partition intersects the requested time range")
    return out


class StoreTrades:
    """Trades loaded from the month store straight into device memory.  Quacks like ``TradesData`` for the bar kits, the
    sigma transforms and ``TBMLabel`` (they all go through ``core.device_trades_for``, which takes the ready device handle);
    ``.data`` materialises the pandas frame lazily from the memory-mapped month files, only when something needs it."""

    def __init__(self, dev: core.DeviceTrades, parts: List[dict], slices: List[slice], has_side: bool, name=None):
        self._dev, self._parts, self._slices, self.has_side, self.name = dev, parts, slices, has_side, name
        self._frame = None

    def __len__(self):
        return self._dev.n

    def device_trades(self, need_ts=False, need_side=False, ctx=None) -> core.DeviceTrades:
        if need_side and not self.has_side:
            raise KeyError('side')
        return self._dev

    def column(self, name) -> np.ndarray:
        """one host column of the loaded range (concatenated from the memory-mapped month files)"""
        arrs = [np.load(os.path.join(p["dir"], name + ".npy"), mmap_mode="r")[s] for p, s in zip(self._parts, self._slices)]
        return np.concatenate(arrs) if len(arrs) != 1 else np.asarray(arrs[0])

    @property
    def data(self) -> pd.DataFrame:
        if self._frame is None:
            ts = self.column("timestamp")
            cols = {"timestamp": ts, "price": self.column("price"), "amount": self.column("amount")}
            if self.has_side:
                cols["side"] = self.column("side")
            df = pd.DataFrame(cols)
            df.set_index(pd.to_datetime(df["timestamp"], unit="ns"), inplace=True)
            df.index.name = "datetime"
            self._frame = df
            core.adopt_device_trades(df, self._dev, has_ts=True, has_side=self.has_side)
        return self._frame


def load_trades_device(root: str, symbol: str, start_time=None, end_time=None, ctx: Optional[core.Context] = None) -> StoreTrades:
    """Load ``[start_time, end_time]`` of a symbol from the month store into ONE device SoA handle (timestamps, price, amount,
    side when stored).  Each month is memory-mapped and copied into the handle at its offset; rows outside the range are cut
    with a binary search on the month's timestamps (the reference filters with a ``where`` clause, data_model.py:640-690)."""
    ctx = ctx or core.default_context()
    parts = list_months(root, symbol, start_time, end_time)
    start = pd.Timestamp(start_time).value if start_time is not None else None
    end = pd.Timestamp(end_time).value if end_time is not None else None
    slices, total = [], 0
    has_side = all(p.get("has_side", False) for p in parts)
    for p in parts:
        ts = np.load(os.path.join(p["dir"], "timestamp.npy"), mmap_mode="r")
        a = int(np.searchsorted(ts, start, "left")) if start is not None else 0
        b = int(np.searchsorted(ts, end, "right")) if end is not None else len(ts)
        slices.append(slice(a, b))
        total += b - a
    if total == 0:
        raise ValueError("no trades inside the requested time range")
    import ctypes as C
    L = ctx._L
    h = C.c_void_p()
    ctx.check(L.fmk_trades_alloc(ctx.h, total, 1, int(has_side), C.byref(h)))
    dev = core.DeviceTrades(ctx, h, total)
    dev.has_ts, dev.has_side = True, has_side
    off = 0
    for p, s in zip(parts, slices):
        n = s.stop - s.start
        if n == 0:
            continue
        cols = {}
        for name, dt in _COLS:
            f = os.path.join(p["dir"], name + ".npy")
            if name == "side" and not has_side:
                continue
            cols[name] = np.load(f, mmap_mode="r")[s]
        am = cols["amount"]
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
        ctx.check(L.fmk_trades_write(ctx.h, h, off, n, ptr(cols["timestamp"]), ptr(cols["price"]), ptr(am),
                                     int(am.dtype == np.float32), ptr(cols["side"]) if has_side else None))
        off += n
    ctx.sync()
    return StoreTrades(dev, parts, slices, has_side, name=symbol)


def add_time_bars(root: str, symbol: str, period: pd.Timedelta = pd.Timedelta(seconds=1), months: Optional[List[str]] = None,
                  overwrite: bool = False, ctx: Optional[core.Context] = None) -> List[str]:
    """``AddTimeBarH5.process_key`` for the columnar store (bar/io.py:441-514): per month, load the trades to the device,
    build ``TimeBarKit(trades, period).build_ohlcv()`` and persist the bars under ``klines/YYYY-MM`` with the reference's
    metadata record (record_count, first / last timestamp, original trades key).  Returns the months processed."""
    from .kit import TimeBarKit
    done = []
    for p in list_months(root, symbol):
        if months is not None and p["key"] not in months:
            continue
        out_dir = os.path.join(root, symbol, "klines", p["key"])
        if os.path.isdir(out_dir) and not overwrite:
            continue
        tr = load_trades_device(root, symbol, pd.Timestamp(p["first_timestamp"], unit="ns"),
                                pd.Timestamp(p["last_timestamp"], unit="ns"), ctx=ctx)
        bars = TimeBarKit(tr, period, ctx=ctx).build_ohlcv()
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, "timestamp.npy"), bars.index.as_unit("ns").asi8)
        for c in bars.columns:
            np.save(os.path.join(out_dir, c + ".npy"), bars[c].values)
        with open(os.path.join(out_dir, "meta.json"), "w") as f:
            json.dump({"record_count": int(len(bars)), "first_timestamp": int(bars.index[0].value), "last_timestamp": int(bars.index[-1].value),
                       "original_trades_key": f"/trades/{p['key']}", "period_seconds": period.total_seconds()}, f)
        done.append(p["key"])
    return done


def load_time_bars(root: str, symbol: str, start_time=None, end_time=None) -> pd.DataFrame:
    """the persisted klines of a time range as one frame (the reading half of ``TimeBarReader``, bar/io.py)"""
    base = os.path.join(root, symbol, "klines")
    frames = []
    for key in sorted(os.listdir(base)):
        d = os.path.join(base, key)
        ts = np.load(os.path.join(d, "timestamp.npy"))
        cols = {f[:-4]: np.load(os.path.join(d, f)) for f in sorted(os.listdir(d)) if f.endswith(".npy") and f != "timestamp.npy"}
        df = pd.DataFrame(cols, index=pd.to_datetime(ts, unit="ns"))
        df.index.name = "timestamp"
        frames.append(df[['open', 'high', 'low', 'close', 'volume', 'trades', 'median_trade_size', 'vwap']])
    out = pd.concat(frames)
    if start_time is not None or end_time is not None:
        out = out.loc[start_time:end_time]
    return out
