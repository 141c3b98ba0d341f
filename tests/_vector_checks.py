"""Shared assertions for the known-answer vectors; `api` is a namespace with the reference's function names."""
import numpy as np
import pytest

import ref_vectors as V


def check_ohlcv(api):
    for c in V.OHLCV:
        o = api.comp_bar_ohlcv(np.array(c["p"]), np.array(c["v"]), np.array(c["idx"], np.int64))
        for k, nm in enumerate(["open", "high", "low", "close"]):
            np.testing.assert_array_equal(o[k], np.array(c[nm]), err_msg=f"{c['name']}.{nm}")
        np.testing.assert_array_equal(o[4], np.array(c["volume"], np.float32), err_msg=c["name"])
        assert o[4].dtype == np.float32 and o[6].dtype == np.int64 and o[0].dtype == np.float64
        np.testing.assert_allclose(o[5], np.array(c["vwap"]), rtol=1e-12, err_msg=c["name"])
        np.testing.assert_array_equal(o[6], np.array(c["trades"]), err_msg=c["name"])
        np.testing.assert_array_equal(o[7], np.array(c["median"]), err_msg=c["name"])
    with pytest.raises(ValueError, match="same length"):
        api.comp_bar_ohlcv(np.array([1., 2.]), np.array([1.]), np.array([0, 1], np.int64))
    with pytest.raises(ValueError, match="at least two"):
        api.comp_bar_ohlcv(np.array([1., 2.]), np.array([1., 2.]), np.array([0], np.int64))


def check_time_clock(api):
    for c in V.TIME_CLOCK:
        clock, idx = api.time_bar_indexer(np.array(c["ts"], np.int64), c["interval"])
        np.testing.assert_array_equal(clock, np.array(c["clock"], np.int64))
        assert len(idx) == len(clock) and np.all(np.diff(idx) >= 0)
        np.testing.assert_array_equal(idx, np.searchsorted(np.array(c["ts"]), np.array(c["clock"]), side="right") - 1)


def check_directional(api):
    for c in V.DIRECTIONAL:
        out = api.comp_bar_directional_features(np.array(c["p"]), np.array(c["v"]), np.array(c["idx"], np.int64), np.array(c["side"], np.int8))
        assert len(out) == 14
        for k, want in enumerate(c["out"]):
            if want is None:
                continue
            if out[k].dtype.kind == "f":
                np.testing.assert_allclose(out[k], np.array(want, np.float32), rtol=1e-6)
            else:
                np.testing.assert_array_equal(out[k], np.array(want))


def check_footprint(api):
    c = V.FOOTPRINT
    r = api.comp_bar_footprints(np.array(c["p"]), np.array(c["a"]), np.array(c["idx"], np.int64), np.array(c["side"], np.int8),
                                c["tick"], np.array(c["lows"]), np.array(c["highs"]), c["factor"])
    assert len(r) == 13 and all(len(x) == 2 for x in r)
    np.testing.assert_array_equal(r[0][0], np.array(c["levels0"], np.int32))
    np.testing.assert_array_equal(r[1][0], np.array(c["buy0"], np.float32))
    np.testing.assert_array_equal(r[2][0], np.array(c["sell0"], np.float32))
    np.testing.assert_array_equal(r[3][0], np.array(c["bt0"], np.int32))
    np.testing.assert_array_equal(r[4][0], np.array(c["st0"], np.int32))
    assert np.asarray(r[0][0]).dtype == np.int32 and np.asarray(r[1][0]).dtype == np.float32
    # bar 1: one sell trade of 2.0 at level 200
    np.testing.assert_array_equal(r[0][1], np.array([200], np.int32))
    np.testing.assert_array_equal(r[2][1], np.array([2.0], np.float32))
    # a trade outside [low, high] must raise like the reference (base.py:719)
    with pytest.raises(ValueError, match="Invalid price level index"):
        api.comp_bar_footprints(np.array(c["p"]), np.array(c["a"]), np.array(c["idx"], np.int64), np.array(c["side"], np.int8),
                                c["tick"], np.array([101.0, 100.0]), np.array(c["highs"]), c["factor"])


def check_tbm(api):
    for c in V.TBM:
        ts = np.arange(len(c["close"]), dtype=np.int64) * V.S
        lab, tch, rets, rat = api.triple_barrier(ts, np.array(c["close"], np.float64), np.array([0], np.int64), np.array([c["tgt"]]),
                                                 (1.0, 1.0), c["vert"], 0.0, None, 0.0)
        assert lab[0] == c["label"] and rat[0] == c["ratio"] and tch[0] < c["touch_lt"]
        assert lab.dtype == np.int8 and tch.dtype == np.int64
    c = V.TBM_TIMEOUT
    ts = np.arange(len(c["close"]), dtype=np.int64) * V.S
    lab, tch, rets, rat = api.triple_barrier(ts, np.array(c["close"], np.float64), np.array([0], np.int64), np.array([c["tgt"]]),
                                             (1.0, 1.0), c["vert"], 0.0, None, 0.0)
    assert tch[0] == c["touch"] and (rat[0] < 1.0 or np.isnan(rat[0]))
    np.testing.assert_allclose(rets[0], np.log(c["close"][c["touch"]] / c["close"][0]), rtol=1e-12)
    for kind, msg in V.TBM_ERRORS:
        with pytest.raises(ValueError, match=msg.replace(".", r"\.")):
            api.triple_barrier(*V.tbm_error_args(kind))


def check_weights(api):
    for c in V.AVG_UNIQUENESS:
        w, conc = api.average_uniqueness(np.arange(c["n"], dtype=np.int64), np.array(c["ev"], np.int64), np.array(c["touch"], np.int64))
        np.testing.assert_allclose(w, np.array(c["w"]), rtol=1e-12)
        np.testing.assert_array_equal(conc, np.array(c["conc"], np.int16))
        assert conc.dtype == np.int16 and w.dtype == np.float64
    c = V.RETURN_ATTRIBUTION
    w = api.return_attribution(np.array(c["ev"], np.int64), np.array(c["touch"], np.int64), np.array(c["close"]), np.array(c["conc"], np.int16), False)
    np.testing.assert_allclose(w, [abs(np.log(102 / 100) + np.log(104 / 102) + np.log(106 / 104))], rtol=1e-12)
    c = V.RETURN_ATTRIBUTION_ZERO
    with pytest.raises(ValueError, match="Sum of weights is zero or negative, cannot normalize"):
        api.return_attribution(np.array(c["ev"], np.int64), np.array(c["touch"], np.int64), np.array(c["close"]), np.array(c["conc"], np.int16), True)
    w = api.return_attribution(np.zeros(0, np.int64), np.zeros(0, np.int64), np.array([100., 101., 102.]), np.ones(3, np.int16), False)
    assert len(w) == 0 and w.dtype == np.float64


def check_volume_profile(vp_rolling_csr):
    """One two-bar series whose second window aggregates exactly the five given levels (no bucketing)."""
    for c in V.VP_ABOVE_POC:
        vol = np.array(c["vol"], np.float32)
        ts = np.array([0, 10**9], np.int64)
        off = np.array([0, 5, 10], np.int64)
        lv = np.tile(np.arange(100, 105, dtype=np.int32), 2)
        buy = np.concatenate([np.zeros(5, np.float32), vol])
        sell = np.zeros(10, np.float32)
        poc, hva, lva, pct = vp_rolling_csr(ts, np.array([10.4, 10.4]), np.array([10.0, 10.0]), off, lv, buy, sell, 0.5, None, 0.1)
        assert poc[0] == 0 and pct[0] == 0.0                  # no full window yet
        assert poc[1] == c["poc"], (poc, c)
        np.testing.assert_allclose(pct[1], np.float32(c["pct"]), rtol=1e-6)
        assert lva[1] <= poc[1] <= hva[1]
