"""GPU parity against the committed reference fixtures (tests/golden/*.npz), through the C ABI."""
import numpy as np
import pytest

from helpers import (STREAM_CASES, assert_exact, assert_f64, check_directional, check_footprint_csr, check_ohlcv,
                     check_trade_size, clock_cases, load_case)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=STREAM_CASES)
def case(request, ctx):
    from finmlkit_b200 import core
    g = load_case(request.param)
    tr = core.DeviceTrades.upload(g["in_ts"], g["in_px"], g["in_qty"], g["in_side"], ctx=ctx)
    return request.param, g, tr


def test_time_bar_indexer(case):
    from finmlkit_b200 import core
    name, g, tr = case
    clock, idx = core.time_bar_index(tr, g["in_params"][0]).download()
    assert_exact(clock, g["ref_time_clock"], f"{name}.time.clock")
    assert_exact(idx, g["ref_time_idx"], f"{name}.time.idx")


@pytest.mark.parametrize("path", clock_cases())
def test_subsecond_clock(path, ctx):
    from finmlkit_b200 import core
    g = dict(np.load(path))
    n = len(g["in_ts"])
    tr = core.DeviceTrades.upload(g["in_ts"], np.ones(n), np.ones(n), ctx=ctx)
    clock, idx = core.time_bar_index(tr, float(g["in_interval"][0])).download()
    assert_exact(clock, g["ref_time_clock"], "clock")
    assert_exact(idx, g["ref_time_idx"], "idx")


def test_tick_bar_indexer(case):
    from finmlkit_b200 import core
    name, g, tr = case
    ts, idx = core.tick_bar_index(tr, int(g["in_params"][1])).download()
    assert_exact(idx, g["ref_tick_idx"], f"{name}.tick.idx")
    assert_exact(ts, g["in_ts"][g["ref_tick_idx"]], f"{name}.tick.ts")


def test_dollar_bar_indexer(case):
    from finmlkit_b200 import core
    name, g, tr = case
    ts, idx = core.dollar_bar_index(tr, g["in_params"][3]).download()
    assert_exact(idx, g["ref_dollar_idx"], f"{name}.dollar.idx")
    assert_exact(ts, g["in_ts"][g["ref_dollar_idx"]], f"{name}.dollar.ts")


def test_volume_bar_indexer(case):
    from finmlkit_b200 import core
    name, g, tr = case
    ts, idx = core.volume_bar_index(tr, g["in_params"][2]).download()
    assert_exact(idx, g["ref_volume_idx"], f"{name}.volume.idx")


def test_cusum_bar_indexer(case):
    from finmlkit_b200 import core
    name, g, tr = case
    sig = core.DeviceBuf.upload(tr.ctx, g["in_cusum_sigma"])
    ts, idx = core.cusum_bar_index(tr, sig, 5e-4, 2.0).download()
    assert_exact(idx, g["ref_cusum_idx"], f"{name}.cusum.idx")
    assert_exact(sig.download(np.float64, len(g["in_ts"])), g["ref_cusum_sigma_filled"], f"{name}.cusum.sigma")


@pytest.mark.parametrize("kind", ["time", "dollar", "volume", "tick", "cusum"])
def test_reductions_on_reference_indices(case, kind):
    """comp_bar_ohlcv / directional / trade-size / footprints keyed on the REFERENCE's close indices."""
    from finmlkit_b200 import core
    name, g, tr = case
    if f"ref_{kind}_idx" not in g or len(g[f"ref_{kind}_idx"]) < 2:
        pytest.skip("no bars")
    what = f"{name}.{kind}"
    ix = core.DeviceIndex.from_host(tr, g[f"ref_{kind}_idx"])
    ref_o = [g[f"ref_{kind}_ohlcv_{n}"] for n in ["open", "high", "low", "close", "volume", "vwap", "trades", "median"]]
    got_o = core.bar_ohlcv(tr, ix)
    check_ohlcv(got_o, ref_o, what)
    check_directional(core.bar_directional(tr, ix), [g[f"ref_{kind}_dir_{k}"] for k in range(14)], what)
    check_trade_size(core.bar_trade_size(tr, ix, g[f"in_{kind}_theta"], 5.0), [g[f"ref_{kind}_ts_{k}"] for k in range(4)], what)
    if f"ref_{kind}_fp_off" in g:
        tick = g["in_params"][4]
        got = core.bar_footprints_csr(tr, ix, tick, ref_o[2], ref_o[1], 3.0)
        scale = float(np.max(np.abs(g[f"ref_{kind}_fp_0"]))) if len(g[f"ref_{kind}_fp_0"]) else 1.0
        check_footprint_csr(got, g[f"ref_{kind}_fp_off"], [g[f"ref_{kind}_fp_{k}"] for k in range(13)], scale, what)


def test_lagged_returns_and_ewmst(case, ctx):
    from finmlkit_b200 import core
    name, g, tr = case
    w, hl = g["in_params"][5], g["in_params"][6]
    r = core.lagged_returns(g["in_ts"], g["in_px"], w, True, ctx=ctx)
    assert_f64(r, g["ref_lagret_log"], f"{name}.lagret.log", rtol=1e-9, atol=1e-15)
    r2 = core.lagged_returns(g["in_ts"], g["in_px"], w, False, ctx=ctx)
    assert_f64(r2, g["ref_lagret_simple"], f"{name}.lagret.simple", rtol=1e-9, atol=1e-15)
    s = core.ewmst_series(g["in_ts"], g["ref_lagret_log"], hl, ctx=ctx)
    assert_f64(s, g["ref_ewmst"], f"{name}.ewmst", rtol=1e-9, atol=1e-18)


def test_triple_barrier(case):
    from finmlkit_b200 import core
    name, g, tr = case
    if "in_tbm_events" not in g:
        pytest.skip("no events")
    ev, tg = g["in_tbm_events"], g["in_tbm_targets"]
    b, t, vert, minc, minret = g["in_tbm_params"]
    lab, tch, rets, rat = core.triple_barrier_dev(tr, ev, tg, (b, t), vert, minc, None, minret)
    skipped = np.isnan(g["ref_tbm_rets"])
    assert_exact(lab, g["ref_tbm_labels"], f"{name}.tbm.labels")
    assert_exact(tch[~skipped], g["ref_tbm_touch"][~skipped], f"{name}.tbm.touch")   # H10: skipped events uninitialised
    assert_f64(rets, g["ref_tbm_rets"], f"{name}.tbm.rets", rtol=1e-9, atol=1e-15)
    assert_f64(rat, g["ref_tbm_ratios"], f"{name}.tbm.ratios", rtol=1e-9, atol=1e-15)
    lab, tch, rets, rat = core.triple_barrier_dev(tr, ev, tg, (1.0, np.inf), vert / 5 * 2, 0.0, g["in_tbm_side"], 1e-4)
    skipped = np.isnan(g["ref_tbm_meta_rets"])
    assert_exact(lab, g["ref_tbm_meta_labels"], f"{name}.tbm.meta.labels")
    assert_exact(tch[~skipped], g["ref_tbm_meta_touch"][~skipped], f"{name}.tbm.meta.touch")
    assert_f64(rets, g["ref_tbm_meta_rets"], f"{name}.tbm.meta.rets", rtol=1e-9, atol=1e-15)
    assert_f64(rat, g["ref_tbm_meta_ratios"], f"{name}.tbm.meta.ratios", rtol=1e-9, atol=1e-15)


def test_bar_level_features(ctx):
    """a15: realized_vol / ewms / vpin / comp_flow_acceleration on the dollar bars of the synthetic fixture."""
    from finmlkit_b200.feature.core.volatility import ewms, realized_vol
    from finmlkit_b200.feature.core.volume import comp_flow_acceleration, vpin
    g = load_case("synth_20k")
    assert_exact(realized_vol(g["in_bar_ret"], 5, True, ctx=ctx), g["ref_rv_5_1"], "rv sample")     # same summation order
    assert_exact(realized_vol(g["in_bar_ret"], 20, False, ctx=ctx), g["ref_rv_20_0"], "rv population")
    assert_f64(ewms(g["in_bar_ret"], 10, ctx=ctx), g["ref_ewms_10"], "ewms", rtol=1e-9, atol=1e-15)
    from helpers import assert_f32_ulp
    assert_f32_ulp(vpin(g["in_vpin_vb"], g["in_vpin_vs"], 8, ctx=ctx), g["ref_vpin_8"], "vpin", ulps=1, atol=1e-9)
    assert_f64(comp_flow_acceleration(g["ref_dollar_ohlcv_volume"].astype(np.float64), 20, 5, ctx=ctx), g["ref_flow_acc_20_5"],
               "flow acc", rtol=1e-9, atol=1e-12)
