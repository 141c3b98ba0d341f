"""
CPU oracle for the finmlkit tick-data hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package.  The product (``finmlkit_b200``) never does.  ``fmk_oracle.c`` restates the reference's Numba functions in
plain C; the functions below give them the reference's own array signatures and return tuples so parity tests read like
the reference's tests (reference paths are cited per function in the C file).

Parity status: PINNED -- checked against fixtures generated from the imported reference (tests/golden/make_golden.py)
and against the reference's hand-written test vectors (tests/test_oracle_golden.py).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfmk_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fmk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.fmko_time_bar_indexer.restype = C.c_int64
        _lib.fmko_tick_bar_indexer.restype = C.c_int64
        _lib.fmko_volume_bar_indexer.restype = C.c_int64
        _lib.fmko_dollar_bar_indexer.restype = C.c_int64
        _lib.fmko_cusum_bar_indexer.restype = C.c_int64
        _lib.fmko_bar_footprints.restype = C.c_int64
        _lib.fmko_cusum_filter.restype = C.c_int64
        _lib.fmko_merge_split_trades.restype = C.c_int64
        _lib.fmko_imbalance_bar_indexer.restype = C.c_int64
        _lib.fmko_imbalance_bar_indexer_ema.restype = C.c_int64
    return _lib


ERRORS = {
    -1: "length mismatch",
    -2: "Bar close indices must contain at least two elements.",
    -3: "capacity",
    -4: "Something went wrong! Invalid price level index!",
    -5: "The vertical barrier must be greater than zero.",
    -6: "The minimum return must be non-negative.",
    -7: "The event_idxs array must not be empty.",
    -8: "The return window must be greater than zero.",
    -9: "Theta should match the the number of bars (len(bar_close_indices) - 1).",
}


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i8(a):
    return np.ascontiguousarray(a, dtype=np.int8)


def num_threads():
    return lib().fmko_num_threads()


def set_num_threads(n):
    lib().fmko_set_num_threads(int(n))


def time_bar_indexer(timestamps, interval_seconds):
    ts = _i64(timestamps)
    n = lib().fmko_time_bar_indexer(_p(ts), C.c_int64(len(ts)), C.c_double(interval_seconds), None, None, C.c_int64(0))
    clock = np.empty(n, np.int64)
    idx = np.empty(n, np.int64)
    lib().fmko_time_bar_indexer(_p(ts), C.c_int64(len(ts)), C.c_double(interval_seconds), _p(clock), _p(idx), C.c_int64(n))
    return clock, idx


def _two_phase(fn, *args):
    n = fn(*args, None, C.c_int64(0))
    idx = np.empty(n, np.int64)
    fn(*args, _p(idx), C.c_int64(n))
    return idx


def tick_bar_indexer(timestamps, threshold):
    return _two_phase(lib().fmko_tick_bar_indexer, C.c_int64(len(timestamps)), C.c_int64(int(threshold)))


def volume_bar_indexer(volumes, threshold):
    v = _f64(volumes)
    return _two_phase(lib().fmko_volume_bar_indexer, _p(v), C.c_int64(len(v)), C.c_double(threshold))


def dollar_bar_indexer(prices, volumes, threshold):
    p, v = _f64(prices), _f64(volumes)
    return _two_phase(lib().fmko_dollar_bar_indexer, _p(p), _p(v), C.c_int64(len(p)), C.c_double(threshold))


def cusum_bar_indexer(timestamps, prices, sigma, sigma_floor, sigma_mult):
    """Forward-fills ``sigma`` in place when it is a contiguous float64 array, like the reference."""
    ts, p = _i64(timestamps), _f64(prices)
    sg = sigma if (isinstance(sigma, np.ndarray) and sigma.dtype == np.float64 and sigma.flags.c_contiguous) else _f64(sigma).copy()
    return _two_phase(lib().fmko_cusum_bar_indexer, _p(ts), _p(p), _p(sg), C.c_int64(len(p)),
                      C.c_double(sigma_floor), C.c_double(sigma_mult))


def comp_bar_ohlcv(prices, volumes, bar_close_indices):
    p, v, ci = _f64(prices), _f64(volumes), _i64(bar_close_indices)
    nb = max(len(ci) - 1, 0)
    o, h, l, c = (np.zeros(nb) for _ in range(4))
    vol = np.zeros(nb, np.float32)
    vwap = np.zeros(nb)
    trades = np.zeros(nb, np.int64)
    med = np.zeros(nb)
    rc = lib().fmko_bar_ohlcv(_p(p), _p(v), C.c_int64(len(p)), C.c_int64(len(v)), _p(ci), C.c_int64(len(ci)),
                              _p(o), _p(h), _p(l), _p(c), _p(vol), _p(vwap), _p(trades), _p(med))
    if rc == -1:
        raise ValueError("Prices and volumes arrays must have the same length.")
    if rc:
        raise ValueError(ERRORS[rc])
    return o, h, l, c, vol, vwap, trades, med


def comp_bar_directional_features(prices, volumes, bar_close_indices, trade_sides):
    p, v, ci, s = _f64(prices), _f64(volumes), _i64(bar_close_indices), _i8(trade_sides)
    nb = len(ci) - 1
    i64 = lambda: np.zeros(nb, np.int64)
    f32 = lambda: np.zeros(nb, np.float32)
    out = [i64(), i64(), f32(), f32(), f32(), f32(), f32(), f32(), i64(), i64(), f32(), f32(), f32(), f32()]
    lib().fmko_bar_directional(_p(p), _p(v), C.c_int64(len(p)), _p(ci), C.c_int64(len(ci)), _p(s), *[_p(a) for a in out])
    return tuple(out)


def comp_bar_trade_size_features(amounts, theta, bar_close_indices, theta_mult):
    a, th, ci = _f64(amounts), _f64(theta), _i64(bar_close_indices)
    nb = len(ci) - 1
    out = [np.zeros(max(nb, 0), np.float32) for _ in range(4)]
    rc = lib().fmko_bar_trade_size(_p(a), C.c_int64(len(a)), _p(th), C.c_int64(len(th)), _p(ci), C.c_int64(len(ci)),
                                   C.c_double(theta_mult), *[_p(x) for x in out])
    if rc:
        raise ValueError(ERRORS[rc])
    return tuple(out)


def comp_bar_footprints_csr(prices, amounts, bar_close_indices, trade_sides, price_tick_size, bar_lows, bar_highs,
                            imbalance_factor):
    """CSR form: (level_offsets, levels, buy_vol, sell_vol, buy_ticks, sell_ticks, buy_imb, sell_imb,
    buy_imb_sum, sell_imb_sum, cot, run_signed, vp_skew, vp_gini)."""
    p, a, ci, s = _f64(prices), _f64(amounts), _i64(bar_close_indices), _i8(trade_sides)
    lo, hi = _f64(bar_lows), _f64(bar_highs)
    nb = len(ci) - 1
    off = np.zeros(nb + 1, np.int64)
    args = [_p(p), _p(a), C.c_int64(len(p)), _p(ci), C.c_int64(len(ci)), _p(s), C.c_double(price_tick_size), _p(lo), _p(hi),
            C.c_double(imbalance_factor), _p(off)]
    total = lib().fmko_bar_footprints(*args, *([None] * 13))
    levels = np.zeros(total, np.int32)
    bv, sv = np.zeros(total, np.float32), np.zeros(total, np.float32)
    bt, st = np.zeros(total, np.int32), np.zeros(total, np.int32)
    bi, si = np.zeros(total, np.bool_), np.zeros(total, np.bool_)
    bis, sis = np.zeros(nb, np.uint16), np.zeros(nb, np.uint16)
    cot = np.zeros(nb, np.int32)
    run = np.zeros(nb, np.int16)
    skew, gini = np.zeros(nb), np.zeros(nb)
    rc = lib().fmko_bar_footprints(*args, _p(levels), _p(bv), _p(sv), _p(bt), _p(st), _p(bi), _p(si), _p(bis), _p(sis),
                                   _p(cot), _p(run), _p(skew), _p(gini))
    if rc < 0:
        raise ValueError(ERRORS[rc])
    return off, levels, bv, sv, bt, st, bi, si, bis, sis, cot, run, skew, gini


def comp_bar_footprints(prices, amounts, bar_close_indices, trade_sides, price_tick_size, bar_lows, bar_highs,
                        imbalance_factor):
    """Reference-shaped return: 7 ragged lists + 6 per-bar arrays (bar/base.py:615-752)."""
    r = comp_bar_footprints_csr(prices, amounts, bar_close_indices, trade_sides, price_tick_size, bar_lows, bar_highs,
                                imbalance_factor)
    off = r[0]
    ragged = [[x[off[i]:off[i + 1]] for i in range(len(off) - 1)] for x in r[1:8]]
    return (*ragged, *r[8:])


def comp_lagged_returns(timestamps, close, return_window_sec, is_log):
    ts, c = _i64(timestamps), _f64(close)
    out = np.empty(len(c))
    rc = lib().fmko_lagged_returns(_p(ts), _p(c), C.c_int64(len(c)), C.c_double(return_window_sec), C.c_int(bool(is_log)), _p(out))
    if rc:
        raise ValueError(ERRORS[rc])
    return out


def ewmst(timestamps, y, half_life, sigma_floor=1e-12):
    ts, yy = _i64(timestamps), _f64(y)
    out = np.empty(len(yy))
    lib().fmko_ewmst(_p(ts), _p(yy), C.c_int64(len(yy)), C.c_double(half_life), C.c_double(sigma_floor), _p(out))
    return out


def triple_barrier(timestamps, close, event_idxs, targets, horizontal_barriers, vertical_barrier, min_close_time_sec,
                   side, min_ret):
    ts, c, ev, tg = _i64(timestamps), _f64(close), _i64(event_idxs), _f64(targets)
    sd = _i8(side) if side is not None else None
    ne = len(ev)
    labels = np.zeros(ne, np.int8)
    touch = np.zeros(ne, np.int64)
    rets = np.full(ne, np.nan)
    ratios = np.full(ne, np.nan)
    bottom, top = horizontal_barriers
    rc = lib().fmko_triple_barrier(_p(ts), _p(c), C.c_int64(len(ts)), C.c_int64(len(c)), _p(ev), _p(tg), C.c_int64(ne),
                                   C.c_int64(len(tg)), C.c_double(bottom), C.c_double(top), C.c_double(vertical_barrier),
                                   C.c_double(min_close_time_sec), _p(sd), C.c_int64(len(sd) if sd is not None else 0),
                                   C.c_double(min_ret), _p(labels), _p(touch), _p(rets), _p(ratios))
    if rc == -1:
        if len(ts) != len(c):
            raise ValueError("The lengths of timestamps and close must match.")
        if ne != len(tg):
            raise ValueError("The lengths of event_idxs and targets must match.")
        raise ValueError("The length of event_idxs must match the length of side.")
    if rc:
        raise ValueError(ERRORS[rc])
    return labels, touch, rets, ratios


def realized_vol(r, window, is_sample):
    rr = _f64(r)
    out = np.empty(len(rr))
    lib().fmko_realized_vol(_p(rr), C.c_int64(len(rr)), C.c_int64(int(window)), C.c_int(bool(is_sample)), _p(out))
    return out


def ewms(y, span):
    yy = _f64(y)
    out = np.empty(len(yy))
    lib().fmko_ewms(_p(yy), C.c_int64(len(yy)), C.c_int64(int(span)), _p(out))
    return out


def vpin(volume_buy, volume_sell, window):
    b, s = _f64(volume_buy), _f64(volume_sell)
    out = np.empty(len(b), np.float32)
    lib().fmko_vpin(_p(b), _p(s), C.c_int64(len(b)), C.c_int64(int(window)), _p(out))
    return out


def comp_flow_acceleration(volumes, window, recent_periods):
    v = _f64(volumes)
    out = np.empty(len(v))
    lib().fmko_flow_acceleration(_p(v), C.c_int64(len(v)), C.c_int64(int(window)), C.c_int64(int(recent_periods)), _p(out))
    return out


def average_uniqueness(timestamps, event_idxs, touch_idxs):
    """label/weights.py:7-49 -> (weights f64[E], concurrency i16[n])."""
    ev, tc = _i64(event_idxs), _i64(touch_idxs)
    n = len(timestamps)
    w = np.zeros(len(ev))
    conc = np.zeros(n, np.int16)
    rc = lib().fmko_average_uniqueness(C.c_int64(n), _p(ev), _p(tc), C.c_int64(len(ev)), C.c_int64(len(tc)), _p(w), _p(conc))
    if rc:
        raise ValueError("Timestamps and lookahead indices must have the same length.")
    return w, conc


def return_attribution(event_idxs, touch_idxs, close, concurrency, normalize):
    """label/weights.py:52-103."""
    ev, tc, c = _i64(event_idxs), _i64(touch_idxs), _f64(close)
    cc = np.ascontiguousarray(concurrency, dtype=np.int16)
    w = np.zeros(len(ev))
    rc = lib().fmko_return_attribution(_p(ev), _p(tc), C.c_int64(len(ev)), _p(c), _p(cc), C.c_int64(len(c)),
                                       C.c_int(bool(normalize)), _p(w))
    if rc == 1:
        raise ValueError("Sum of weights is zero or negative, cannot normalize.")
    return w


def cusum_filter(raw_time_series, threshold):
    """sampling/filters.py:6-70 -> event indices int64."""
    x, t = _f64(raw_time_series), _f64(threshold)
    args = (_p(x), C.c_int64(len(x)), _p(t), C.c_int64(len(t)))
    m = lib().fmko_cusum_filter(*args, None, C.c_int64(0))
    if m == -1:
        raise ValueError("Input time series must have at least 2 elements.")
    if m == -2:
        raise ValueError("Threshold array must either contain 1 const. element or len(raw_time_series) elements.")
    out = np.empty(m, np.int64)
    lib().fmko_cusum_filter(*args, _p(out), C.c_int64(m))
    return out


def comp_trade_side_vector(prices):
    """bar/utils.py:26-46 (tick rule) -> int8 sides."""
    p = _f64(prices)
    out = np.zeros(len(p), np.int8)
    lib().fmko_trade_side_vector(_p(p), C.c_int64(len(p)), _p(out))
    return out


def merge_split_trades(timestamps, prices, amounts, is_buyer_maker):
    """bar/utils.py:263-329 -> (ts i64, price f64, amount f32, side i8 or empty)."""
    ts, p = _i64(timestamps), _f64(prices)
    a = np.ascontiguousarray(amounts, dtype=np.float32)
    ibm = np.ascontiguousarray(is_buyer_maker, dtype=np.uint8) if is_buyer_maker is not None else None
    n = len(ts)
    ots, op, oa = np.empty(n, np.int64), np.empty(n), np.empty(n, np.float32)
    osd = np.empty(n, np.int8) if ibm is not None else None
    m = lib().fmko_merge_split_trades(_p(ts), _p(p), _p(a), _p(ibm), C.c_int64(n), _p(ots), _p(op), _p(oa), _p(osd))
    return ots[:m], op[:m], oa[:m], (osd[:m] if osd is not None else np.empty(0, np.int8))


def volume_profile_rolling_csr(ts, highs, lows, level_offsets, price_levels, buy_volumes, sell_volumes, window_size_sec,
                               n_bins, price_tick, va_pct=68.34):
    """feature/core/volume.py:396-456 on a CSR footprint -> (poc i32, hva i32, lva i32, vp_pct_abv_poc f32)."""
    t, h, l = _i64(ts), _f64(highs), _f64(lows)
    off = _i64(level_offsets)
    lv = np.ascontiguousarray(price_levels, dtype=np.int32)
    b, s_ = np.ascontiguousarray(buy_volumes, dtype=np.float32), np.ascontiguousarray(sell_volumes, dtype=np.float32)
    nb = len(t)
    poc, hva, lva = (np.zeros(nb, np.int32) for _ in range(3))
    pct = np.zeros(nb, np.float32)
    lib().fmko_volume_profile_rolling(_p(t), _p(h), _p(l), C.c_int64(nb), _p(off), _p(lv), _p(b), _p(s_), C.c_double(window_size_sec),
                                      C.c_int64(int(n_bins) if n_bins else 0), C.c_double(price_tick), C.c_double(va_pct),
                                      _p(poc), _p(hva), _p(lva), _p(pct))
    return poc, hva, lva, pct


# ---- a6: tick-imbalance / tick-run bars: OWN semantics, parity UNPINNED (the reference only has stubs, logic.py:224-261) ----
def imbalance_bar_indexer(sides, threshold, kind=0):
    """kind 0: |running sum of b_t| >= threshold closes; kind 1: max(#buys, #sells) >= threshold closes.  idx[0] = 0."""
    b = _i8(sides)
    return _two_phase(lib().fmko_imbalance_bar_indexer, _p(b), C.c_int64(len(b)), C.c_double(float(threshold)), C.c_int(int(kind)))


def imbalance_bar_indexer_ema(sides, expected_ticks_init, expected_imbalance_init, span_bars, thr_min, thr_max):
    """EMA-adaptive tick-imbalance bars (AFML 2.3.2.1, ewma in the adjust=True form of ma.py:7-43) -> (idx, thresholds)."""
    b = _i8(sides)
    args = (_p(b), C.c_int64(len(b)), C.c_double(float(expected_ticks_init)), C.c_double(float(expected_imbalance_init)),
            C.c_int64(int(span_bars)), C.c_double(float(thr_min)), C.c_double(float(thr_max)))
    m = lib().fmko_imbalance_bar_indexer_ema(*args, None, None, C.c_int64(0))
    idx = np.zeros(m, np.int64)
    thr = np.full(m, np.nan)
    lib().fmko_imbalance_bar_indexer_ema(*args, _p(idx), _p(thr), C.c_int64(m))
    return idx, thr
