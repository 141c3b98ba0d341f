"""
Thin object layer over the C ABI (include/fmk.h): Context, DeviceTrades, DeviceIndex and the array-level functions that
mirror the reference's Numba "core functions" (same argument meaning, same return tuples, same ValueError messages).

Everything here runs on the GPU through libfmk.so; nothing falls back to the CPU.
"""
import ctypes as C
import weakref

import numpy as np

from ._lib import lib

_ERRORS = {-1: "CUDA", -2: "ARG", -3: "ALLOC", -4: "CAPACITY", -5: "LEVEL", -6: "INTERNAL"}


class FmkError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Context:
    """One CUDA device + stream (fmk_ctx). Single-threaded; create one per device."""

    def __init__(self, device: int = 0, stream: int = None):
        self._L = lib()
        if self._L.fmk_device_count() <= 0:
            raise FmkError("no CUDA device: finmlkit_b200 has no CPU fallback")
        h = C.c_void_p()
        if stream is None:
            rc = self._L.fmk_ctx_create(int(device), C.byref(h))
        else:
            rc = self._L.fmk_ctx_create_on_stream(int(device), C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise FmkError(f"fmk_ctx_create(device={device}) failed: {_ERRORS.get(rc, rc)}")
        self.h = h
        self.device = device
        self._fin = weakref.finalize(self, self._L.fmk_ctx_destroy, h)

    def check(self, rc):
        if rc == 0:
            return
        msg = self._L.fmk_last_error(self.h).decode()
        if rc in (-2, -5):
            raise ValueError(msg)          # the reference raises ValueError with the same text
        raise FmkError(f"{_ERRORS.get(rc, rc)}: {msg}")

    def sync(self):
        self.check(self._L.fmk_ctx_sync(self.h))

    def timer_start(self):
        self.check(self._L.fmk_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.check(self._L.fmk_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self._L.fmk_launch_count(self.h))

    def prof_enable(self, on=True):
        self._L.fmk_prof_enable(self.h, int(on))

    def prof_report(self, cap=64):
        """{kernel name: (launches, total ms)} of everything launched since the last report while profiling was on."""
        names = C.create_string_buffer(64 * cap)
        counts = np.zeros(cap, np.int64)
        ms = np.zeros(cap, np.float32)
        nk = self._L.fmk_prof_report(self.h, names, _ptr(counts), _ptr(ms), cap)
        out = {}
        for k in range(nk):
            nm = names.raw[64 * k:64 * (k + 1)].split(b"\0", 1)[0].decode()
            out[nm] = (int(counts[k]), float(ms[k]))
        return out

    def result_cols(self):
        p, nb, nbytes = C.c_void_p(), C.c_int64(), C.c_int64()
        self._L.fmk_result_cols(self.h, C.byref(p), C.byref(nb), C.byref(nbytes))
        return p.value, int(nb.value), int(nbytes.value)

    def flush_l2(self):
        self.check(self._L.fmk_flush_l2(self.h))

    def index_stats(self):
        s = np.zeros(3, np.int64)
        self._L.fmk_index_stats(self.h, _ptr(s))
        return {"tasks": int(s[0]), "serial_repairs": int(s[1]), "chain_passes": int(s[2])}


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceBuf:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        self._fin = weakref.finalize(self, ctx._L.fmk_buf_free, ctx.h, h)

    @classmethod
    def upload(cls, ctx, arr):
        arr = np.ascontiguousarray(arr)
        h = C.c_void_p()
        ctx.check(ctx._L.fmk_buf_upload(ctx.h, _ptr(arr), arr.nbytes, C.byref(h)))
        ctx.sync()
        return cls(ctx, h)

    def download(self, dtype, count):
        out = np.empty(count, dtype)
        self.ctx.check(self.ctx._L.fmk_buf_download(self.ctx.h, self.h, _ptr(out), out.nbytes))
        return out

    def gather(self, idx, dtype=np.float64):
        """src[idx] for host indices into a device array of 8-byte elements, without downloading the array."""
        ii = _c(idx, np.int64)
        out = np.empty(len(ii), dtype)
        assert out.itemsize == 8
        self.ctx.check(self.ctx._L.fmk_buf_gather8(self.ctx.h, self.h, _ptr(ii), len(ii), _ptr(out)))
        return out

    @property
    def devptr(self):
        return self.ctx._L.fmk_buf_devptr(self.h)

    @property
    def nbytes(self):
        return int(self.ctx._L.fmk_buf_bytes(self.h))


class DeviceTrades:
    """Trade columns as device SoA (fmk_trades): ts int64 ns, price f64, amount f64, side int8 (optional)."""

    def __init__(self, ctx: Context, h, n):
        self.ctx, self.h, self.n = ctx, h, n
        self.has_ts = True
        self._fin = weakref.finalize(self, ctx._L.fmk_trades_free, ctx.h, h)

    @classmethod
    def upload(cls, ts, price, amount, side=None, ctx: Context = None):
        ctx = ctx or default_context()
        # float32 amounts (TradesData after the reference's split-trade merge, data_model.py:326-344) cross PCIe as 4 B/tick
        # and are widened exactly on the device; anything else is float64
        f32 = isinstance(amount, np.ndarray) and amount.dtype == np.float32
        price, amount = _c(price, np.float64), _c(amount, np.float32 if f32 else np.float64)
        ts = _c(ts, np.int64) if ts is not None else None      # None: the host keeps the timestamps (see fmk.h)
        if len(price) != len(amount) or (ts is not None and len(ts) != len(price)):
            raise ValueError("Prices and volumes arrays must have the same length.")
        sd = _c(side, np.int8) if side is not None else None
        h = C.c_void_p()
        up = ctx._L.fmk_trades_upload_f32amt if f32 else ctx._L.fmk_trades_upload
        ctx.check(up(ctx.h, _ptr(ts), _ptr(price), _ptr(amount), _ptr(sd), len(price), C.byref(h)))
        ctx.sync()
        obj = cls(ctx, h, len(price))
        obj.has_ts = ts is not None
        return obj

    @classmethod
    def synth(cls, n, seed=42, ctx: Context = None):
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx._L.fmk_trades_synth(ctx.h, int(n), int(seed), C.byref(h)))
        ctx.sync()
        return cls(ctx, h, int(n))

    def refill(self, ts, price, amount, side=None):
        """Async H2D of new column contents into this handle (arrays must stay alive until the next sync)."""
        self.ctx.check(self.ctx._L.fmk_trades_refill(self.ctx.h, self.h, _ptr(ts), _ptr(price), _ptr(amount), _ptr(side), self.n))

    def download(self, out=None):
        if out is None:
            out = (np.empty(self.n, np.int64), np.empty(self.n, np.float64), np.empty(self.n, np.float64), np.empty(self.n, np.int8))
        ts, px, qty, side = out
        self.ctx.check(self.ctx._L.fmk_trades_download(self.ctx.h, self.h, _ptr(ts), _ptr(px), _ptr(qty), _ptr(side)))
        return ts, px, qty, side


# ---- one upload per TradesData frame, shared by every kit / transform / label call that is handed the same frame -------------
_TRADES_CACHE = {}      # id(frame) -> (weakref to the frame, price buffer address, n, ctx, DeviceTrades, has_side)
_SERIES_CACHE = {}      # id(ndarray) -> (weakref to the array, DeviceBuf): device copies of tick-length series (returns, sigma)


def device_trades_for(df, need_ts=False, need_side=False, ctx: Context = None) -> DeviceTrades:
    """The device SoA copy of a trades frame (``TradesData.data``: columns timestamp, price, amount[, side]), uploaded ONCE
    and shared by every builder, transform and label call that receives the same frame object; the timestamp and side columns
    are attached lazily, when a caller first needs them.  The entry dies with the frame (weak reference)."""
    ctx = ctx or default_context()
    px = df['price'].values
    key = id(df)
    hit = _TRADES_CACHE.get(key)
    if hit is not None:
        ref, addr, n, hctx, tr, _ = hit
        if ref() is df and addr == px.ctypes.data and n == len(px) and hctx is ctx:
            pass
        else:
            hit = None
    if hit is None:
        tr = DeviceTrades.upload(None, px, df['amount'].values, None, ctx=ctx)
        tr.has_ts, tr.has_side = False, False

        def _drop(_r, key=key):
            _TRADES_CACHE.pop(key, None)
        _TRADES_CACHE[key] = (weakref.ref(df, _drop), px.ctypes.data, len(px), ctx, tr, False)
    tr = _TRADES_CACHE[key][4]
    if need_ts and not tr.has_ts:
        ts = _c(df['timestamp'].values.astype(np.int64, copy=False), np.int64)
        ctx.check(ctx._L.fmk_trades_add_column(ctx.h, tr.h, 0, _ptr(ts)))
        ctx.sync()
        tr.has_ts = True
    if need_side and not getattr(tr, "has_side", False):
        if 'side' not in df.columns:
            raise KeyError('side')
        sd = _c(df['side'].values, np.int8)
        ctx.check(ctx._L.fmk_trades_add_column(ctx.h, tr.h, 1, _ptr(sd)))
        ctx.sync()
        tr.has_side = True
    return tr


def adopt_device_trades(df, tr: DeviceTrades, has_ts=True, has_side=False):
    """Register an existing device handle as THE device copy of the frame ``df`` (bar.io.StoreTrades: the columns were loaded
    from the month store straight into the device, the pandas frame was built from the same files afterwards)."""
    px = df['price'].values
    key = id(df)
    tr.has_ts, tr.has_side = has_ts, has_side

    def _drop(_r, key=key):
        _TRADES_CACHE.pop(key, None)
    _TRADES_CACHE[key] = (weakref.ref(df, _drop), px.ctypes.data, len(px), tr.ctx, tr, has_side)


def _root_array(arr):
    """the ndarray that owns the memory ``arr`` views (pandas hands out a fresh read-only view per ``Series.values`` call)"""
    while isinstance(getattr(arr, "base", None), np.ndarray):
        arr = arr.base
    return arr


def register_device_copy(arr, payload):
    """Remember that ``payload`` (a DeviceBuf, or (DeviceBuf, DeviceTrades)) is the device copy of the host series ``arr`` --
    e.g. the sigma a fused Compose returned: a later call that is handed the same memory (CUSUMBarKit(trades, sigma), EWMST on
    the returns) skips the upload.  ``arr`` may be a Series: the entry is keyed on the array that OWNS the memory and dies with
    it (weak reference), so recycled memory can never alias a stale entry.  In-place edits by the caller are not detected."""
    vals = arr.values if hasattr(arr, "values") and not isinstance(arr, np.ndarray) else arr
    if not isinstance(vals, np.ndarray):
        return
    root = _root_array(vals)
    key = id(root)

    def _drop(_r, key=key):
        _SERIES_CACHE.pop(key, None)
    try:
        _SERIES_CACHE[key] = (weakref.ref(root, _drop), vals.ctypes.data, vals.nbytes, payload)
    except TypeError:
        pass


def device_copy_of(arr):
    if not isinstance(arr, np.ndarray):
        return None
    root = _root_array(arr)
    hit = _SERIES_CACHE.get(id(root))
    if hit is not None and hit[0]() is root and hit[1] == arr.ctypes.data and hit[2] == arr.nbytes:
        return hit[3]
    return None


def clear_device_cache():
    _TRADES_CACHE.clear()
    _SERIES_CACHE.clear()


class DeviceIndex:
    """Bar close timestamps / indices on the device (fmk_index), m = n_bars + 1 elements."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        self.m = int(ctx._L.fmk_index_size(h))
        self._fin = weakref.finalize(self, ctx._L.fmk_index_free, ctx.h, h)

    def download(self, host_ts=None, out_idx=None, gather=True):
        """(close_ts, close_idx).  When the trades were uploaded without timestamps pass the host array: close_ts is
        then ts[close_idx] gathered on the host (bar/kit.py:67 `close_ts = timestamps[close_indices]`).
        ``out_idx``: optional preallocated int64 buffer of >= m elements (e.g. pinned memory, for repeated calls);
        ``gather=False`` returns (None, close_idx) so that the caller can overlap the host gather with GPU work."""
        idx = np.empty(self.m, np.int64) if out_idx is None else out_idx[:self.m]
        ts = np.empty(self.m, np.int64) if host_ts is None else None
        self.ctx.check(self.ctx._L.fmk_index_download(self.ctx.h, self.h, _ptr(ts), _ptr(idx)))
        if host_ts is not None and gather:
            ts = np.asarray(host_ts)[idx]
        return ts, idx

    @classmethod
    def from_host(cls, trades: DeviceTrades, close_idx):
        ci = _c(close_idx, np.int64)
        h = C.c_void_p()
        ctx = trades.ctx
        ctx.check(ctx._L.fmk_index_from_host(ctx.h, trades.h, _ptr(ci), len(ci), C.byref(h)))
        ctx.sync()
        return cls(ctx, h)


def _mk_index(trades: DeviceTrades, fn, *args):
    h = C.c_void_p()
    trades.ctx.check(fn(trades.ctx.h, trades.h, *args, C.byref(h)))
    return DeviceIndex(trades.ctx, h)


def time_bar_index(trades: DeviceTrades, interval_seconds: float) -> DeviceIndex:
    return _mk_index(trades, trades.ctx._L.fmk_time_bar_index, float(interval_seconds))


def tick_bar_index(trades: DeviceTrades, threshold: int) -> DeviceIndex:
    return _mk_index(trades, trades.ctx._L.fmk_tick_bar_index, int(threshold))


def volume_bar_index(trades: DeviceTrades, threshold: float) -> DeviceIndex:
    return _mk_index(trades, trades.ctx._L.fmk_volume_bar_index, float(threshold))


def dollar_bar_index(trades: DeviceTrades, threshold: float) -> DeviceIndex:
    return _mk_index(trades, trades.ctx._L.fmk_dollar_bar_index, float(threshold))


def cusum_bar_index(trades: DeviceTrades, sigma: DeviceBuf, sigma_floor: float, sigma_mult: float) -> DeviceIndex:
    return _mk_index(trades, trades.ctx._L.fmk_cusum_bar_index, sigma.h, float(sigma_floor), float(sigma_mult))


def imbalance_bar_index(trades: DeviceTrades, threshold: float, use_side: bool = True, kind: int = 0) -> DeviceIndex:
    """tick-imbalance (kind 0) / tick-run (kind 1) bars -- own semantics, parity unpinned (include/fmk.h)"""
    return _mk_index(trades, trades.ctx._L.fmk_imbalance_bar_index, float(threshold), int(bool(use_side)), int(kind))


# ---- per-bar reductions ---------------------------------------------------------------------------------------------
def bar_ohlcv(trades: DeviceTrades, index: DeviceIndex, median=True, out=None):
    """(open, high, low, close, volume f32, vwap, trades i64, median) -- tuple order of comp_bar_ohlcv (base.py:407).
    ``out``: optional tuple of 8 preallocated arrays of >= n_bars elements in that order (e.g. pinned memory)."""
    nb = max(index.m - 1, 0)
    if out is not None:
        o, h, l, c, vol, vwap, tr, med = (a[:nb] for a in out)
        if not median:
            med = None
    else:
        o, h, l, c, vwap = (np.empty(nb) for _ in range(5))
        vol = np.empty(nb, np.float32)
        tr = np.empty(nb, np.int64)
        med = np.empty(nb) if median else None
    ctx = trades.ctx
    ctx.check(ctx._L.fmk_bar_ohlcv(ctx.h, trades.h, index.h, _ptr(o), _ptr(h), _ptr(l), _ptr(c), _ptr(vol), _ptr(vwap), _ptr(tr), _ptr(med)))
    return o, h, l, c, vol, vwap, tr, med


def bar_directional(trades: DeviceTrades, index: DeviceIndex):
    nb = max(index.m - 1, 0)
    i64 = lambda: np.empty(nb, np.int64)   # noqa: E731
    f32 = lambda: np.empty(nb, np.float32)  # noqa: E731
    out = [i64(), i64(), f32(), f32(), f32(), f32(), f32(), f32(), i64(), i64(), f32(), f32(), f32(), f32()]
    ctx = trades.ctx
    ctx.check(ctx._L.fmk_bar_directional(ctx.h, trades.h, index.h, *[_ptr(a) for a in out]))
    return tuple(out)


def bar_trade_size(trades: DeviceTrades, index: DeviceIndex, theta, theta_mult):
    nb = max(index.m - 1, 0)
    th = _c(theta, np.float64)
    out = [np.empty(nb, np.float32) for _ in range(4)]
    ctx = trades.ctx
    ctx.check(ctx._L.fmk_bar_trade_size(ctx.h, trades.h, index.h, _ptr(th), len(th), float(theta_mult), *[_ptr(a) for a in out]))
    return tuple(out)


def bar_footprints_csr(trades: DeviceTrades, index: DeviceIndex, price_tick_size, bar_lows, bar_highs, imbalance_factor):
    """CSR footprint: (level_offsets, levels, buy_vol, sell_vol, buy_ticks, sell_ticks, buy_imb, sell_imb,
    buy_imb_sum, sell_imb_sum, cot, run_signed, vp_skew, vp_gini)."""
    nb = max(index.m - 1, 0)
    lo, hi = _c(bar_lows, np.float64), _c(bar_highs, np.float64)
    if len(lo) != nb or len(hi) != nb:
        raise ValueError("bar_lows / bar_highs must have one element per bar")
    ctx = trades.ctx
    h = C.c_void_p()
    ctx.check(ctx._L.fmk_bar_footprints(ctx.h, trades.h, index.h, float(price_tick_size), _ptr(lo), _ptr(hi), float(imbalance_factor), C.byref(h)))
    try:
        nl = int(ctx._L.fmk_footprint_levels(h))
        off = np.empty(nb + 1, np.int64)
        levels = np.empty(nl, np.int32)
        bv, sv = np.empty(nl, np.float32), np.empty(nl, np.float32)
        bt, st = np.empty(nl, np.int32), np.empty(nl, np.int32)
        bi, si = np.empty(nl, np.bool_), np.empty(nl, np.bool_)
        bis, sis = np.empty(nb, np.uint16), np.empty(nb, np.uint16)
        cot = np.empty(nb, np.int32)
        run = np.empty(nb, np.int16)
        skew, gini = np.empty(nb), np.empty(nb)
        ctx.check(ctx._L.fmk_footprint_download(ctx.h, h, _ptr(off), _ptr(levels), _ptr(bv), _ptr(sv), _ptr(bt), _ptr(st), _ptr(bi),
                                                _ptr(si), _ptr(bis), _ptr(sis), _ptr(cot), _ptr(run), _ptr(skew), _ptr(gini)))
    finally:
        ctx._L.fmk_footprint_free(ctx.h, h)
    return off, levels, bv, sv, bt, st, bi, si, bis, sis, cot, run, skew, gini


# ---- tick-level series -----------------------------------------------------------------------------------------------
def lagged_returns(timestamps, close, return_window_sec, is_log, ctx: Context = None):
    ctx = ctx or default_context()
    if return_window_sec <= 0:
        raise ValueError("The return window must be greater than zero.")
    ts, c = _c(timestamps, np.int64), _c(close, np.float64)
    out = np.empty(len(c))
    ctx.check(ctx._L.fmk_lagged_returns(ctx.h, _ptr(ts), _ptr(c), len(c), float(return_window_sec), int(bool(is_log)), _ptr(out)))
    return out


def ewmst_series(timestamps, y, half_life, sigma_floor=1e-12, ctx: Context = None):
    ctx = ctx or default_context()
    ts, yy = _c(timestamps, np.int64), _c(y, np.float64)
    out = np.empty(len(yy))
    ctx.check(ctx._L.fmk_ewmst(ctx.h, _ptr(ts), _ptr(yy), len(yy), float(half_life), float(sigma_floor), _ptr(out)))
    return out


def lagged_returns_dev(trades: DeviceTrades, return_window_sec, is_log) -> DeviceBuf:
    """comp_lagged_returns on the device-resident ts / price columns; the result stays on the device."""
    if return_window_sec <= 0:
        raise ValueError("The return window must be greater than zero.")
    h = C.c_void_p()
    ctx = trades.ctx
    ctx.check(ctx._L.fmk_lagged_returns_dev(ctx.h, trades.h, float(return_window_sec), int(bool(is_log)), C.byref(h)))
    return DeviceBuf(ctx, h)


def ewmst_dev(trades: DeviceTrades, y: DeviceBuf, half_life, sigma_floor=1e-12) -> DeviceBuf:
    """ewmst of a device-resident series on the trades' timestamps; the result stays on the device."""
    h = C.c_void_p()
    ctx = trades.ctx
    ctx.check(ctx._L.fmk_ewmst_dev(ctx.h, trades.h, y.h, float(half_life), float(sigma_floor), C.byref(h)))
    return DeviceBuf(ctx, h)


def triple_barrier_dev(trades: DeviceTrades, event_idxs, targets, horizontal_barriers, vertical_barrier, min_close_time_sec, side, min_ret):
    ev, tg = _c(event_idxs, np.int64), _c(targets, np.float64)
    sd = _c(side, np.int8) if side is not None else None
    ne = len(ev)
    labels = np.zeros(ne, np.int8)
    touch = np.zeros(ne, np.int64)
    rets = np.full(ne, np.nan)
    ratios = np.full(ne, np.nan)
    bottom, top = horizontal_barriers
    ctx = trades.ctx
    ctx.check(ctx._L.fmk_triple_barrier(ctx.h, trades.h, _ptr(ev), _ptr(tg), ne, len(tg), float(bottom), float(top), float(vertical_barrier),
                                        float(min_close_time_sec), _ptr(sd), len(sd) if sd is not None else 0, float(min_ret),
                                        _ptr(labels), _ptr(touch), _ptr(rets), _ptr(ratios)))
    return labels, touch, rets, ratios


# ---- bar-level features (n_bars-length arrays) -----------------------------------------------------------------------
def realized_vol_series(r, window, is_sample, ctx: Context = None):
    ctx = ctx or default_context()
    rr = _c(r, np.float64)
    out = np.empty(len(rr))
    ctx.check(ctx._L.fmk_realized_vol(ctx.h, _ptr(rr), len(rr), int(window), int(bool(is_sample)), _ptr(out)))
    return out


def ewms_series(y, span, ctx: Context = None):
    ctx = ctx or default_context()
    yy = _c(y, np.float64)
    out = np.empty(len(yy))
    ctx.check(ctx._L.fmk_ewms(ctx.h, _ptr(yy), len(yy), int(span), _ptr(out)))
    return out


def vpin_series(volume_buy, volume_sell, window, ctx: Context = None):
    ctx = ctx or default_context()
    b, s = _c(volume_buy, np.float64), _c(volume_sell, np.float64)
    out = np.empty(len(b), np.float32)
    ctx.check(ctx._L.fmk_vpin(ctx.h, _ptr(b), _ptr(s), len(b), int(window), _ptr(out)))
    return out


def flow_acceleration_series(volumes, window, recent_periods, ctx: Context = None):
    ctx = ctx or default_context()
    v = _c(volumes, np.float64)
    out = np.empty(len(v))
    ctx.check(ctx._L.fmk_flow_acceleration(ctx.h, _ptr(v), len(v), int(window), int(recent_periods), _ptr(out)))
    return out


# ---- sample weights on ticks (label/weights.py) ----------------------------------------------------------------------
def average_uniqueness_dev(n, event_idxs, touch_idxs, ctx: Context = None):
    ctx = ctx or default_context()
    ev, tc = _c(event_idxs, np.int64), _c(touch_idxs, np.int64)
    w = np.zeros(len(ev))
    conc = np.zeros(int(n), np.int16)
    ctx.check(ctx._L.fmk_average_uniqueness(ctx.h, int(n), _ptr(ev), _ptr(tc), len(ev), len(tc), _ptr(w), _ptr(conc)))
    return w, conc


def return_attribution_dev(event_idxs, touch_idxs, close, concurrency, normalize, ctx: Context = None):
    ctx = ctx or default_context()
    ev, tc = _c(event_idxs, np.int64), _c(touch_idxs, np.int64)
    c, cc = _c(close, np.float64), _c(concurrency, np.int16)
    if len(cc) != len(c) or len(ev) != len(tc):
        raise ValueError("close / concurrency and event / touch arrays must have matching lengths")
    w = np.zeros(len(ev))
    ctx.check(ctx._L.fmk_return_attribution(ctx.h, _ptr(ev), _ptr(tc), len(ev), _ptr(c), _ptr(cc), len(c), int(bool(normalize)), _ptr(w)))
    return w


def sample_weights_dev(trades: DeviceTrades, event_idxs, touch_idxs, normalize=False, want_concurrency=False):
    """(avg_uniqueness, return_attribution[, concurrency]) in one pass over the device-resident price column."""
    ev, tc = _c(event_idxs, np.int64), _c(touch_idxs, np.int64)
    if len(ev) != len(tc):
        raise ValueError("Timestamps and lookahead indices must have the same length.")
    u, r = np.zeros(len(ev)), np.zeros(len(ev))
    conc = np.zeros(trades.n, np.int16) if want_concurrency else None
    ctx = trades.ctx
    ctx.check(ctx._L.fmk_sample_weights(ctx.h, trades.h, _ptr(ev), _ptr(tc), len(ev), int(bool(normalize)), _ptr(u), _ptr(r), _ptr(conc)))
    return (u, r, conc) if want_concurrency else (u, r)


# ---- event sampler and ingest scans ------------------------------------------------------------------------------------
def cusum_filter_dev(raw_time_series, threshold, ctx: Context = None):
    ctx = ctx or default_context()
    x, th = _c(raw_time_series, np.float64), _c(threshold, np.float64)
    h, m = C.c_void_p(), C.c_int64()
    ctx.check(ctx._L.fmk_cusum_filter(ctx.h, _ptr(x), len(x), _ptr(th), len(th), C.byref(h), C.byref(m)))
    return DeviceBuf(ctx, h).download(np.int64, int(m.value))


def trade_side_vector_dev(prices, ctx: Context = None):
    ctx = ctx or default_context()
    p = _c(prices, np.float64)
    out = np.zeros(len(p), np.int8)
    ctx.check(ctx._L.fmk_trade_side_vector(ctx.h, _ptr(p), len(p), _ptr(out)))
    return out


def merge_split_trades_dev(timestamps, prices, amounts, is_buyer_maker, ctx: Context = None):
    ctx = ctx or default_context()
    ts, p, a = _c(timestamps, np.int64), _c(prices, np.float64), _c(amounts, np.float32)
    if not (len(ts) == len(p) == len(a)):
        raise ValueError("timestamps, prices and amounts must have the same length")
    ibm = _c(is_buyer_maker, np.uint8) if is_buyer_maker is not None else None
    n = len(ts)
    ots, op, oa = np.empty(n, np.int64), np.empty(n), np.empty(n, np.float32)
    osd = np.empty(n, np.int8) if ibm is not None else None
    m = C.c_int64()
    ctx.check(ctx._L.fmk_merge_split_trades(ctx.h, _ptr(ts), _ptr(p), _ptr(a), _ptr(ibm), n, _ptr(ots), _ptr(op), _ptr(oa), _ptr(osd), C.byref(m)))
    m = int(m.value)
    return ots[:m], op[:m], oa[:m], (osd[:m] if osd is not None else np.empty(0, np.int8))


# ---- rolling volume profile (feature/core/volume.py) --------------------------------------------------------------------
def volume_profile_rolling_csr(ts, highs, lows, level_offsets, price_levels, buy_volumes, sell_volumes, window_size_sec,
                               n_bins, price_tick, va_pct=68.34, ctx: Context = None):
    """(poc i32, hva i32, lva i32, vp_pct_abv_poc f32) from a host CSR footprint."""
    ctx = ctx or default_context()
    t, h, l = _c(ts, np.int64), _c(highs, np.float64), _c(lows, np.float64)
    off, lv = _c(level_offsets, np.int64), _c(price_levels, np.int32)
    b, s = _c(buy_volumes, np.float32), _c(sell_volumes, np.float32)
    nb = len(t)
    if not (len(h) == len(l) == nb == len(off) - 1) or nb == 0:
        raise AssertionError("Input arrays should have the same length and be non-empty.")
    poc, hva, lva = (np.zeros(nb, np.int32) for _ in range(3))
    pct = np.zeros(nb, np.float32)
    ctx.check(ctx._L.fmk_volume_profile_rolling(ctx.h, _ptr(off), _ptr(lv), _ptr(b), _ptr(s), nb, _ptr(t), _ptr(h), _ptr(l),
                                                float(window_size_sec), int(n_bins) if n_bins else 0, float(price_tick),
                                                float(va_pct), _ptr(poc), _ptr(hva), _ptr(lva), _ptr(pct)))
    return poc, hva, lva, pct


# ---- device-resident bar frame (fmk_bar_features_device) ---------------------------------------------------------------
F_OHLCV, F_MEDIAN, F_DIRECTIONAL, F_TRADE_SIZE, F_FOOTPRINT = 1, 2, 4, 8, 16
F_ALL = F_OHLCV | F_MEDIAN | F_DIRECTIONAL | F_TRADE_SIZE | F_FOOTPRINT

# column id -> (name, dtype, block, extra elements); order = the fmk_col enum of include/fmk.h
FRAME_COLS = [
    ("close_ts", np.int64, "bar", 0), ("close_idx", np.int64, "bar", 0),
    ("open", np.float64, "bar", 0), ("high", np.float64, "bar", 0), ("low", np.float64, "bar", 0), ("close", np.float64, "bar", 0),
    ("vwap", np.float64, "bar", 0), ("median_trade_size", np.float64, "bar", 0), ("trades", np.int64, "bar", 0),
    ("volume", np.float32, "bar", 0),
    ("ticks_buy", np.int64, "bar", 0), ("ticks_sell", np.int64, "bar", 0), ("cum_ticks_min", np.int64, "bar", 0),
    ("cum_ticks_max", np.int64, "bar", 0),
    ("volume_buy", np.float32, "bar", 0), ("volume_sell", np.float32, "bar", 0), ("dollars_buy", np.float32, "bar", 0),
    ("dollars_sell", np.float32, "bar", 0), ("mean_spread", np.float32, "bar", 0), ("max_spread", np.float32, "bar", 0),
    ("cum_volume_min", np.float32, "bar", 0), ("cum_volume_max", np.float32, "bar", 0), ("cum_dollars_min", np.float32, "bar", 0),
    ("cum_dollars_max", np.float32, "bar", 0),
    ("mean_size_rel", np.float32, "bar", 0), ("size_95_rel", np.float32, "bar", 0), ("pct_block", np.float32, "bar", 0),
    ("size_gini", np.float32, "bar", 0),
    ("fp_level_offsets", np.int64, "bar", 1), ("fp_vp_skew", np.float64, "bar", 0), ("fp_vp_gini", np.float64, "bar", 0),
    ("fp_cot", np.int32, "bar", 0), ("fp_buy_imb_sum", np.uint16, "bar", 0), ("fp_sell_imb_sum", np.uint16, "bar", 0),
    ("fp_run_signed", np.int16, "bar", 0),
    ("fp_price_levels", np.int32, "level", 0), ("fp_buy_vol", np.float32, "level", 0), ("fp_sell_vol", np.float32, "level", 0),
    ("fp_buy_ticks", np.int32, "level", 0), ("fp_sell_ticks", np.int32, "level", 0), ("fp_buy_imb", np.bool_, "level", 0),
    ("fp_sell_imb", np.bool_, "level", 0),
]


def frame_views(bar_block: np.ndarray, level_block, n_bars, n_levels, col_offsets):
    """{column name: NumPy view} into host copies of a frame's two blocks (uint8 arrays)."""
    out = {}
    for k, (name, dt, blk, extra) in enumerate(FRAME_COLS):
        off = int(col_offsets[k])
        if off < 0:
            continue
        src, cnt = (bar_block, n_bars + extra) if blk == "bar" else (level_block, n_levels)
        if src is None:
            continue
        out[name] = np.frombuffer(src, dtype=dt, count=cnt, offset=off)
    return out


FRAME_HEADER_BYTES = 512
FRAME_MAGIC = 0x464D4B4652414D45


def frame_views_from_bytes(buf) -> dict:
    """{column: array} from the raw bytes of a frame as it crosses the wire (fmk_comm_gather_submit of
    ``DeviceFrame.segments()``: the self-describing per-bar block, then the per-level block at the next 16-byte boundary)."""
    b = np.ascontiguousarray(buf, dtype=np.uint8)
    hdr = np.frombuffer(b, np.int64, FRAME_HEADER_BYTES // 8)
    if int(hdr[0]) != FRAME_MAGIC:
        raise ValueError("not a finmlkit_b200 bar frame")
    nb, nl, bar_bytes, lvl_bytes, ncols = int(hdr[1]), int(hdr[2]), int(hdr[3]), int(hdr[4]), int(hdr[6])
    lvl0 = (bar_bytes + 15) // 16 * 16
    cols = np.full(len(FRAME_COLS), -1, np.int64)
    cols[:min(ncols, len(FRAME_COLS))] = hdr[8:8 + min(ncols, len(FRAME_COLS))]
    return frame_views(b[:bar_bytes], b[lvl0:lvl0 + lvl_bytes] if lvl_bytes else None, nb, nl, cols)


class DeviceFrame:
    """Every per-bar output of one index on the device (fmk_frame): two blocks, described by ``col_offsets``."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        nb, nl, bb, lb = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self.col_offsets = np.zeros(len(FRAME_COLS), np.int64)
        ctx._L.fmk_frame_info(h, C.byref(nb), C.byref(nl), C.byref(bb), C.byref(lb), _ptr(self.col_offsets))
        self.n_bars, self.n_levels, self.bar_bytes, self.level_bytes = int(nb.value), int(nl.value), int(bb.value), int(lb.value)
        self._fin = weakref.finalize(self, ctx._L.fmk_frame_free, ctx.h, h)

    def devptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        self.ctx._L.fmk_frame_devptrs(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def segments(self):
        """[(device pointer, bytes)] of the non-empty blocks -- what fmk_comm_gather_submit takes."""
        a, b = self.devptrs()
        return [(p, n) for p, n in ((a, self.bar_bytes), (b, self.level_bytes)) if n > 0]

    def download(self, bar_out=None, level_out=None):
        """{column: array}; ``bar_out`` / ``level_out``: optional preallocated uint8 buffers (e.g. pinned)."""
        bar = np.empty(self.bar_bytes, np.uint8) if bar_out is None else bar_out[:self.bar_bytes]
        lvl = None
        if self.level_bytes > 0:
            lvl = np.empty(self.level_bytes, np.uint8) if level_out is None else level_out[:self.level_bytes]
        self.ctx.check(self.ctx._L.fmk_frame_download(self.ctx.h, self.h, _ptr(bar), _ptr(lvl)))
        return frame_views(bar, lvl, self.n_bars, self.n_levels, self.col_offsets)

    def free(self):
        self._fin()


def bar_features_device(trades: DeviceTrades, index: DeviceIndex, flags=F_OHLCV | F_MEDIAN, theta=None, theta_mult=5.0,
                        price_tick_size=0.0, imbalance_factor=3.0) -> DeviceFrame:
    th = _c(theta, np.float64) if theta is not None else None
    h = C.c_void_p()
    ctx = trades.ctx
    ctx.check(ctx._L.fmk_bar_features_device(ctx.h, trades.h, index.h, int(flags), _ptr(th), len(th) if th is not None else 0,
                                             float(theta_mult), float(price_tick_size), float(imbalance_factor), C.byref(h)))
    return DeviceFrame(ctx, h)
