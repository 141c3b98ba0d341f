"""Month-partitioned columnar store (the reference's /trades/YYYY-MM HDF5 layout, data_model.py:420-574).
CPU: partitioning, metadata and range discovery.  GPU: the loader fills ONE device handle that equals a plain upload of the
same rows, the kits accept it in place of a TradesData, and add_time_bars persists the klines the reference's AddTimeBarH5
would (bar/io.py:441-514)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from finmlkit_b200.bar import io as fio


def _three_months(n=30_000, seed=3):
    rng = np.random.default_rng(seed)
    t0 = pd.Timestamp("2024-01-20").value
    t1 = pd.Timestamp("2024-03-10").value
    ts = np.sort(rng.integers(t0, t1, n)).astype(np.int64) // 1_000_000 * 1_000_000
    px = np.round(100 + np.cumsum(rng.normal(0, 0.05, n)), 1)
    qty = np.round(rng.lognormal(-2, 1, n) + 0.001, 3).astype(np.float32)
    side = rng.choice([-1, 1], n).astype(np.int8)
    return ts, px, qty, side


def test_save_partitions_by_month_and_discovers_ranges(tmp_path):
    ts, px, qty, side = _three_months()
    keys = fio.save_trades_store(str(tmp_path), "BTCUSDT", ts, px, qty, side)
    assert keys == ["2024-01", "2024-02", "2024-03"]
    months = fio.list_months(str(tmp_path), "BTCUSDT")
    assert [m["key"] for m in months] == keys and sum(m["record_count"] for m in months) == len(ts)
    for m in months:
        t = np.load(os.path.join(m["dir"], "timestamp.npy"))
        assert m["first_timestamp"] == t[0] and m["last_timestamp"] == t[-1] and m["amount_dtype"] == "float32" and m["has_side"]
        assert all(pd.Timestamp(int(x), unit="ns").strftime("%Y-%m") == m["key"] for x in (t[0], t[-1]))
        assert json.load(open(os.path.join(m["dir"], "meta.json")))["record_count"] == len(t)
    feb = fio.list_months(str(tmp_path), "BTCUSDT", "2024-02-03", "2024-02-20")
    assert [m["key"] for m in feb] == ["2024-02"]
    span = fio.list_months(str(tmp_path), "BTCUSDT", "2024-01-31 23:00", "2024-02-01 01:00")
    assert [m["key"] for m in span] == ["2024-01", "2024-02"]
    with pytest.raises(KeyError):
        fio.list_months(str(tmp_path), "BTCUSDT", "2025-01-01", "2025-02-01")
    with pytest.raises(KeyError):
        fio.list_months(str(tmp_path), "ETHUSDT")
    with pytest.raises(ValueError):
        fio.save_trades_store(str(tmp_path), "X", ts[::-1], px, qty)


@pytest.mark.gpu
def test_loader_fills_one_device_handle(tmp_path, ctx):
    import oracle
    from finmlkit_b200 import core
    from finmlkit_b200.bar.data_model import TradesData
    from finmlkit_b200.bar.kit import DollarBarKit, TimeBarKit, VolumeBarKit
    from helpers import assert_exact, check_directional, check_ohlcv
    ts, px, qty, side = _three_months(60_000)
    fio.save_trades_store(str(tmp_path), "BTCUSDT", ts, px, qty, side)
    st = fio.load_trades_device(str(tmp_path), "BTCUSDT", ctx=ctx)
    back = st.device_trades().download()
    assert_exact(back[0], ts, "ts"); assert_exact(back[1], px, "px"); assert_exact(back[3], side, "side")
    assert_exact(back[2], qty.astype(np.float64), "float32 amounts widened on the device")
    # a sub-range that cuts two months
    a, b = pd.Timestamp("2024-01-29 12:00"), pd.Timestamp("2024-02-17 06:30")
    sub = fio.load_trades_device(str(tmp_path), "BTCUSDT", a, b, ctx=ctx)
    sel = (ts >= a.value) & (ts <= b.value)
    assert len(sub) == int(sel.sum())
    assert_exact(sub.device_trades().download()[0], ts[sel], "range ts")
    # the kits take the store object where they take a TradesData, without building the pandas frame
    q64 = qty.astype(np.float64)
    k = DollarBarKit(st, 2e4, ctx=ctx)
    df = k.build_ohlcv()
    ref = oracle.dollar_bar_indexer(px, q64, 2e4)
    assert_exact(k.bar_close_indices, ref[1:], "dollar idx from the store")
    assert_exact(df.index.as_unit("ns").asi8, ts[ref[1:]], "close ts gathered on the device")
    check_ohlcv([df[c].values for c in ("open", "high", "low", "close", "volume", "vwap", "trades", "median_trade_size")],
                oracle.comp_bar_ohlcv(px, q64, ref), "store ohlcv")
    d = VolumeBarKit(st, 30.0, ctx=ctx).build_directional_features()
    vref = oracle.volume_bar_indexer(q64, 30.0)
    check_directional([d[c].values for c in d.columns], oracle.comp_bar_directional_features(px, q64, vref, side), "store directional")
    assert st._frame is None                                            # nothing above needed pandas
    # the lazily built frame matches and is adopted as the frame's device copy (no second upload)
    td = TradesData(ts, px, qty, side=side)
    pd.testing.assert_frame_equal(st.data[["timestamp", "price", "amount", "side"]], td.data[["timestamp", "price", "amount", "side"]])
    assert core.device_trades_for(st.data, need_ts=True, need_side=True, ctx=ctx) is st.device_trades()
    # AddTimeBarH5.process_key equivalent: 1-second klines per month, persisted with the reference's metadata
    done = fio.add_time_bars(str(tmp_path), "BTCUSDT", pd.Timedelta(seconds=1), ctx=ctx)
    assert done == ["2024-01", "2024-02", "2024-03"]
    assert fio.add_time_bars(str(tmp_path), "BTCUSDT", ctx=ctx) == []   # already there, overwrite=False
    kl = fio.load_time_bars(str(tmp_path), "BTCUSDT")
    m = [x for x in fio.list_months(str(tmp_path), "BTCUSDT")]
    feb = (ts >= m[1]["first_timestamp"]) & (ts <= m[1]["last_timestamp"])
    exp = TimeBarKit(TradesData(ts[feb], px[feb], qty[feb], side=side[feb]), pd.Timedelta(seconds=1), ctx=ctx).build_ohlcv()
    got = kl.loc[exp.index[0]:exp.index[-1]]
    pd.testing.assert_frame_equal(got.reset_index(drop=True), exp.reset_index(drop=True))
    meta = json.load(open(os.path.join(str(tmp_path), "BTCUSDT", "klines", "2024-02", "meta.json")))
    assert meta["record_count"] == len(exp) and meta["original_trades_key"] == "/trades/2024-02"
