"""CPU: the C oracle against the committed reference fixtures (pins oracle/ to the reference's outputs)."""
import numpy as np
import pytest

import oracle
from helpers import (STREAM_CASES, assert_exact, check_footprint_csr, clock_cases, load_case)


@pytest.fixture(scope="module", params=STREAM_CASES)
def case(request):
    return request.param, load_case(request.param)


def test_indexers(case):
    name, g = case
    clock, idx = oracle.time_bar_indexer(g["in_ts"], g["in_params"][0])
    assert_exact(clock, g["ref_time_clock"], "clock")
    assert_exact(idx, g["ref_time_idx"], "time idx")
    assert_exact(oracle.tick_bar_indexer(g["in_ts"], int(g["in_params"][1])), g["ref_tick_idx"], "tick")
    assert_exact(oracle.volume_bar_indexer(g["in_qty"], g["in_params"][2]), g["ref_volume_idx"], "volume")
    assert_exact(oracle.dollar_bar_indexer(g["in_px"], g["in_qty"], g["in_params"][3]), g["ref_dollar_idx"], "dollar")
    sig = g["in_cusum_sigma"].copy()
    assert_exact(oracle.cusum_bar_indexer(g["in_ts"], g["in_px"], sig, 5e-4, 2.0), g["ref_cusum_idx"], "cusum")
    assert_exact(sig, g["ref_cusum_sigma_filled"], "cusum sigma in-place forward fill")


@pytest.mark.parametrize("path", clock_cases())
def test_subsecond_clock(path):
    g = dict(np.load(path))
    clock, idx = oracle.time_bar_indexer(g["in_ts"], float(g["in_interval"][0]))
    assert_exact(clock, g["ref_time_clock"], "clock")
    assert_exact(idx, g["ref_time_idx"], "idx")


@pytest.mark.parametrize("kind", ["time", "dollar", "volume", "tick", "cusum"])
def test_reductions(case, kind):
    name, g = case
    if f"ref_{kind}_idx" not in g or len(g[f"ref_{kind}_idx"]) < 2:
        pytest.skip("no bars")
    idx = g[f"ref_{kind}_idx"]
    px, qty, side = g["in_px"], g["in_qty"], g["in_side"]
    o = oracle.comp_bar_ohlcv(px, qty, idx)
    for k, n in enumerate(["open", "high", "low", "close", "volume", "vwap", "trades", "median"]):
        assert_exact(o[k], g[f"ref_{kind}_ohlcv_{n}"], f"ohlcv.{n}")
    d = oracle.comp_bar_directional_features(px, qty, idx, side)
    for k in range(14):
        assert_exact(d[k], g[f"ref_{kind}_dir_{k}"], f"dir[{k}]")
    t = oracle.comp_bar_trade_size_features(qty, g[f"in_{kind}_theta"], idx, 5.0)
    for k in range(4):
        assert_exact(t[k], g[f"ref_{kind}_ts_{k}"], f"tsize[{k}]")
    if f"ref_{kind}_fp_off" in g:
        got = oracle.comp_bar_footprints_csr(px, qty, idx, side, g["in_params"][4], o[2], o[1], 3.0)
        scale = float(np.max(np.abs(g[f"ref_{kind}_fp_0"])))
        check_footprint_csr(got, g[f"ref_{kind}_fp_off"], [g[f"ref_{kind}_fp_{k}"] for k in range(13)], scale, kind)


def test_series_and_labels(case):
    name, g = case
    w, hl = g["in_params"][5], g["in_params"][6]
    assert_exact(oracle.comp_lagged_returns(g["in_ts"], g["in_px"], w, True), g["ref_lagret_log"], "lagret log")
    assert_exact(oracle.comp_lagged_returns(g["in_ts"], g["in_px"], w, False), g["ref_lagret_simple"], "lagret simple")
    assert_exact(oracle.ewmst(g["in_ts"], g["ref_lagret_log"], hl), g["ref_ewmst"], "ewmst")
    if "in_tbm_events" in g:
        b, t, vert, minc, minret = g["in_tbm_params"]
        r = oracle.triple_barrier(g["in_ts"], g["in_px"], g["in_tbm_events"], g["in_tbm_targets"], (b, t), vert, minc, None, minret)
        sk = np.isnan(g["ref_tbm_rets"])
        assert_exact(r[0], g["ref_tbm_labels"], "labels")
        assert_exact(r[1][~sk], g["ref_tbm_touch"][~sk], "touch")
        assert_exact(r[2], g["ref_tbm_rets"], "rets")
        assert_exact(r[3], g["ref_tbm_ratios"], "ratios")
        r = oracle.triple_barrier(g["in_ts"], g["in_px"], g["in_tbm_events"], g["in_tbm_targets"], (1.0, np.inf), vert / 5 * 2, 0.0, g["in_tbm_side"], 1e-4)
        sk = np.isnan(g["ref_tbm_meta_rets"])
        assert_exact(r[0], g["ref_tbm_meta_labels"], "meta labels")
        assert_exact(r[1][~sk], g["ref_tbm_meta_touch"][~sk], "meta touch")
        assert_exact(r[2], g["ref_tbm_meta_rets"], "meta rets")
        assert_exact(r[3], g["ref_tbm_meta_ratios"], "meta ratios")


def test_bar_level_features():
    g = load_case("synth_20k")
    assert_exact(oracle.realized_vol(g["in_bar_ret"], 5, True), g["ref_rv_5_1"], "rv")
    assert_exact(oracle.realized_vol(g["in_bar_ret"], 20, False), g["ref_rv_20_0"], "rv pop")
    assert_exact(oracle.ewms(g["in_bar_ret"], 10), g["ref_ewms_10"], "ewms")
    assert_exact(oracle.vpin(g["in_vpin_vb"], g["in_vpin_vs"], 8), g["ref_vpin_8"], "vpin")
    assert_exact(oracle.comp_flow_acceleration(g["ref_dollar_ohlcv_volume"].astype(np.float64), 20, 5), g["ref_flow_acc_20_5"], "flow")
