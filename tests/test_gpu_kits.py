"""GPU: the pandas-in / pandas-out wrapper layer (kits, TBMLabel, transforms) against the golden fixtures."""
import numpy as np
import pandas as pd
import pytest

from helpers import assert_exact, assert_f64, check_directional, check_ohlcv, check_trade_size, load_case

pytestmark = pytest.mark.gpu

OHLCV_COLS = ['open', 'high', 'low', 'close', 'volume', 'trades', 'median_trade_size', 'vwap']
DIR_COLS = ['ticks_buy', 'ticks_sell', 'volume_buy', 'volume_sell', 'dollars_buy', 'dollars_sell', 'mean_spread', 'max_spread',
            'cum_ticks_min', 'cum_ticks_max', 'cum_volume_min', 'cum_volume_max', 'cum_dollars_min', 'cum_dollars_max']


@pytest.fixture(scope="module")
def g():
    return load_case("synth_20k")


@pytest.fixture(scope="module")
def trades(g):
    from finmlkit_b200.bar.data_model import TradesData
    return TradesData(g["in_ts"], g["in_px"], g["in_qty"], side=g["in_side"])


def _ref_ohlcv(g, kind):
    return [g[f"ref_{kind}_ohlcv_{n}"] for n in ["open", "high", "low", "close", "volume", "vwap", "trades", "median"]]


def _frame_to_tuple(df):
    return [df[c].values for c in ['open', 'high', 'low', 'close', 'volume', 'vwap', 'trades', 'median_trade_size']]


@pytest.mark.parametrize("kind", ["time", "tick", "volume", "dollar", "cusum"])
def test_kit_build_ohlcv(g, trades, kind):
    from finmlkit_b200.bar import kit
    p = g["in_params"]
    k = {"time": lambda: kit.TimeBarKit(trades, pd.Timedelta(seconds=p[0])), "tick": lambda: kit.TickBarKit(trades, int(p[1])),
         "volume": lambda: kit.VolumeBarKit(trades, p[2]), "dollar": lambda: kit.DollarBarKit(trades, p[3]),
         "cusum": lambda: kit.CUSUMBarKit(trades, g["in_cusum_sigma"].copy(), 5e-4, 2.0)}[kind]()
    df = k.build_ohlcv()
    ref_idx = g[f"ref_{kind}_idx"]
    assert list(df.columns) == OHLCV_COLS and df.index.name == "timestamp"
    assert df['volume'].dtype == np.float32 and df['trades'].dtype == np.int64
    assert_exact(k.bar_close_indices, ref_idx[1:], f"{kind}.bar_close_indices")
    ref_ts = g["ref_time_clock"][1:] if kind == "time" else g["in_ts"][ref_idx[1:]]
    assert_exact(df.index.as_unit("ns").asi8, ref_ts, f"{kind}.index")
    assert_exact(k.bar_close_timestamps, ref_ts, f"{kind}.bar_close_timestamps")
    if kind == "time":
        assert df.index.freq == pd.Timedelta(seconds=p[0])
    check_ohlcv(_frame_to_tuple(df), _ref_ohlcv(g, kind), kind)
    d = k.build_directional_features()
    assert list(d.columns) == DIR_COLS
    check_directional([d[c].values for c in DIR_COLS], [g[f"ref_{kind}_dir_{i}"] for i in range(14)], kind)
    t = k.build_trade_size_features(g[f"in_{kind}_theta"], 5.0)
    assert list(t.columns) == ['mean_size_rel', 'size_95_rel', 'pct_block', 'size_gini']
    check_trade_size([t[c].values for c in t.columns], [g[f"ref_{kind}_ts_{i}"] for i in range(4)], kind)
    if kind == "cusum":
        assert_exact(k.get_sigma(), g["ref_cusum_sigma_filled"][ref_idx[1:]], "get_sigma")


def test_kit_build_footprints(g, trades):
    from finmlkit_b200.bar import kit
    k = kit.DollarBarKit(trades, g["in_params"][3])
    fp = k.build_footprints(price_tick_size=None, imbalance_factor=3.0)     # tick size inferred on the host (0.1)
    assert fp.price_tick == pytest.approx(0.1, rel=1e-9) and fp.is_valid() and len(fp) == len(g["ref_dollar_idx"]) - 1
    off = g["ref_dollar_fp_off"]
    for name, kf in [("price_levels", 0), ("buy_volumes", 1), ("sell_volumes", 2), ("buy_ticks", 3), ("sell_ticks", 4),
                     ("buy_imbalances", 5), ("sell_imbalances", 6)]:
        got = getattr(fp, name)
        flat = np.concatenate([np.asarray(x) for x in got])
        assert_exact(flat, g[f"ref_dollar_fp_{kf}"], name)
        assert len(got[3]) == off[4] - off[3]
    assert_exact(fp.buy_imbalances_sum, g["ref_dollar_fp_7"], "buy_imb_sum")
    assert_exact(fp.cot_price_levels, g["ref_dollar_fp_9"], "cot")
    assert_exact(fp.imb_max_run_signed, g["ref_dollar_fp_10"], "run")
    sl = fp[2:5]
    assert len(sl) == 3 and np.array_equal(np.asarray(sl.price_levels[0]), np.asarray(fp.price_levels[2]))
    df = fp.get_df()
    assert list(df.columns) == ['price_level', 'sell_ticks', 'buy_ticks', 'sell_volume', 'buy_volume', 'sell_imbalance', 'buy_imbalance']
    assert df.index.names == ['bar_idx', 'bar_datetime_idx'] and len(df) == off[-1]


def test_mock_builder_with_host_indices(g, trades):
    """A subclass that returns host-computed indices (the reference's MockBarBuilder pattern) still works."""
    from finmlkit_b200.bar.base import BarBuilderBase

    class Mock(BarBuilderBase):
        def _comp_bar_close(self):
            idx = g["ref_tick_idx"]
            return g["in_ts"][idx], idx

    check_ohlcv(_frame_to_tuple(Mock(trades).build_ohlcv()), _ref_ohlcv(g, "tick"), "mock")


def test_sigma_pipeline_and_tbm_label(g, trades):
    from finmlkit_b200.feature.transforms import EWMST, Compose, ReturnT
    from finmlkit_b200.label.kit import TBMLabel
    w, hl = g["in_params"][5], g["in_params"][6]
    sig = Compose(ReturnT(pd.Timedelta(seconds=w), is_log=True, input_col="price"), EWMST(pd.Timedelta(seconds=hl)))(trades.data)
    assert sig.name == f"price_ret{w}s_ewms{hl}s"
    assert_f64(sig.values, g["ref_ewmst"], "sigma", rtol=1e-9, atol=1e-18)
    with pytest.raises(ValueError):
        ReturnT(pd.Timedelta(seconds=w), input_col="price")(trades.data, backend="cuda")
    ev = g["in_tbm_events"]
    feats = pd.DataFrame({"sigma": g["in_tbm_targets"], "event_idx": ev}, index=pd.to_datetime(g["in_ts"][ev], unit="ns"))
    b, t, vert, minc, minret = g["in_tbm_params"]
    lab = TBMLabel(feats, "sigma", min_ret=0.0, horizontal_barriers=(b, t), vertical_barrier=pd.Timedelta(seconds=vert),
                   min_close_time=pd.Timedelta(seconds=minc))
    f, out = lab.compute_labels(trades)
    assert list(out.columns) == ['touch_time', 'event_idx', 'touch_idx', 'labels', 'returns', 'vertical_touch_weights']
    n = len(out)
    assert n > 0 and n <= len(ev)
    assert_exact(out['labels'].values, g["ref_tbm_labels"][:n], "labels")
    assert_exact(out['touch_idx'].values, g["ref_tbm_touch"][:n], "touch")
    assert_f64(out['returns'].values, g["ref_tbm_rets"][:n], "returns", atol=1e-15)


def test_one_upload_per_trades_frame(g, trades, monkeypatch):
    """kits, the sigma transforms, CUSUM bars, TBM labels and sample weights on the SAME TradesData share
    one device copy of the frame -- the stream is uploaded once, and sigma never goes back up."""
    from finmlkit_b200 import core
    from finmlkit_b200.bar import kit
    from finmlkit_b200.feature.transforms import EWMST, Compose, ReturnT
    from finmlkit_b200.label.kit import TBMLabel
    core.clear_device_cache()
    uploads = {"trades": 0, "buf": 0}
    real_t, real_b = core.DeviceTrades.upload.__func__, core.DeviceBuf.upload.__func__
    monkeypatch.setattr(core.DeviceTrades, "upload", classmethod(lambda cls, *a, **k: (uploads.__setitem__("trades", uploads["trades"] + 1), real_t(cls, *a, **k))[1]))
    monkeypatch.setattr(core.DeviceBuf, "upload", classmethod(lambda cls, *a, **k: (uploads.__setitem__("buf", uploads["buf"] + 1), real_b(cls, *a, **k))[1]))
    p = g["in_params"]
    k1 = kit.DollarBarKit(trades, p[3])
    check_ohlcv(_frame_to_tuple(k1.build_ohlcv()), _ref_ohlcv(g, "dollar"), "dollar")
    k2 = kit.TimeBarKit(trades, pd.Timedelta(seconds=p[0]))                  # needs the timestamps: attached, not re-uploaded
    check_ohlcv(_frame_to_tuple(k2.build_ohlcv()), _ref_ohlcv(g, "time"), "time")
    d = k1.build_directional_features()                                      # needs the side column: attached lazily
    check_directional([d[c].values for c in DIR_COLS], [g[f"ref_dollar_dir_{i}"] for i in range(14)], "dollar")
    assert k1._device() is k2._device()
    w, hl = p[5], p[6]
    sig = Compose(ReturnT(pd.Timedelta(seconds=w), is_log=True, input_col="price"), EWMST(pd.Timedelta(seconds=hl)))(trades.data)
    assert_f64(sig.values, g["ref_ewmst"], "fused sigma", rtol=1e-9, atol=1e-18)
    # the separate calls give the same series (returns come back to the host here, sigma is computed from their device copy)
    r = ReturnT(pd.Timedelta(seconds=w), is_log=True, input_col="price")(trades.data)
    s2 = EWMST(pd.Timedelta(seconds=hl), input_col=r.name)(r.to_frame())
    assert_exact(s2.values, sig.values, "unfused == fused")
    assert uploads == {"trades": 1, "buf": 0}, uploads
    ck = kit.CUSUMBarKit(trades, sig.values, 5e-4, 2.0)                       # sigma's device copy is reused
    idx = ck.bar_close_indices
    assert uploads == {"trades": 1, "buf": 0}, uploads
    import oracle
    assert_exact(idx, oracle.cusum_bar_indexer(g["in_ts"], g["in_px"], sig.values.copy(), 5e-4, 2.0)[1:], "cusum on device sigma")
    ev = idx[np.isfinite(sig.values[idx])]
    feats = pd.DataFrame({"sigma": sig.values[ev], "event_idx": ev}, index=pd.to_datetime(g["in_ts"][ev], unit="ns"))
    lab = TBMLabel(feats, "sigma", min_ret=0.0, horizontal_barriers=(2.0, 2.0), vertical_barrier=pd.Timedelta(seconds=600))
    _, out = lab.compute_labels(trades)
    wts = lab.compute_weights(trades)
    assert len(wts) == len(out) and uploads["trades"] == 1, uploads
    core.clear_device_cache()


def test_transforms_subclass_the_reference_when_it_is_importable(g):
    """With the reference on sys.path (baseline/_ref) ReturnT / EWMST / Compose ARE the reference's CoreTransform classes with
    the GPU behind _nb: they can sit inside a reference Feature.  Run in a subprocess: the import order matters."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "finmlkit")):
        pytest.skip("baseline/_ref not installed")
    code = r'''
import sys, numpy as np, pandas as pd
sys.path[:0] = [%r, %r, %r]
from helpers import load_case
import finmlkit.feature.transforms as rt, finmlkit.feature.kit as rk
from finmlkit.feature.kit import Feature
from finmlkit_b200.feature.transforms import ReturnT, EWMST, Compose
from finmlkit_b200.bar.data_model import TradesData
assert issubclass(ReturnT, rt.ReturnT) and issubclass(EWMST, rt.EWMST) and issubclass(Compose, rk.Compose)
g = load_case("synth_20k")
td = TradesData(g["in_ts"], g["in_px"], g["in_qty"], side=g["in_side"])
w, hl = g["in_params"][5], g["in_params"][6]
c = Compose(ReturnT(pd.Timedelta(seconds=w), is_log=True, input_col="price"), EWMST(pd.Timedelta(seconds=hl)))
sig = c(td.data)
assert sig.name == f"price_ret{w}s_ewms{hl}s", sig.name
assert np.allclose(sig.values, g["ref_ewmst"], rtol=1e-9, atol=1e-18, equal_nan=True)
f = Feature(c)(td.data)                       # the reference's Feature wrapper around the GPU-backed Compose
assert np.allclose(f.values, g["ref_ewmst"], rtol=1e-9, atol=1e-18, equal_nan=True)
try:
    ReturnT(pd.Timedelta(seconds=w), input_col="price")(td.data, backend="cuda")
    raise SystemExit("no ValueError for an unknown backend")
except ValueError:
    pass
print("SUBCLASS_OK")
''' % (os.path.join(root, "tests"), root, ref)
    env = dict(os.environ, FMK_CONSOLE_LOGGER_LEVEL="ERROR")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SUBCLASS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
