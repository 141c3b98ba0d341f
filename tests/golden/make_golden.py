"""
Generate golden fixtures from the UNMODIFIED reference (quantscious/finmlkit, imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    FMK_CONSOLE_LOGGER_LEVEL=ERROR PYTHONPATH=/root/reference python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Each file holds the inputs (``in_*``) and the reference outputs (``ref_*``) of one case.
Footprint ragged lists are stored in CSR form (``ref_fp_off`` + flat arrays).  The script also cross-checks the C oracle
against the reference on 1M-tick streams (not stored) and prints the result.
"""
import os
import sys

os.environ.setdefault("FMK_CONSOLE_LOGGER_LEVEL", "ERROR")
sys.path.insert(0, "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import numpy as np  # noqa: E402

from finmlkit.bar.logic import (_time_bar_indexer, _tick_bar_indexer, _volume_bar_indexer, _dollar_bar_indexer,  # noqa: E402
                                _cusum_bar_indexer)
from finmlkit.bar.base import (comp_bar_ohlcv, comp_bar_directional_features, comp_bar_trade_size_features,  # noqa: E402
                               comp_bar_footprints)
from finmlkit.feature.core.utils import comp_lagged_returns  # noqa: E402
from finmlkit.feature.core.volatility import ewmst, ewms, realized_vol  # noqa: E402
from finmlkit.feature.core.volume import vpin, comp_flow_acceleration  # noqa: E402
from finmlkit.label.tbm import triple_barrier  # noqa: E402

from finmlkit_b200.synth import synth_trades  # noqa: E402


def csr(lists, dtype):
    off = np.zeros(len(lists) + 1, np.int64)
    for i, x in enumerate(lists):
        off[i + 1] = off[i] + len(x)
    flat = np.concatenate([np.asarray(x) for x in lists]).astype(dtype) if len(lists) else np.zeros(0, dtype)
    return off, flat


def ref_bundle(ts, px, qty, side, idx, prefix, out, tick=0.1, theta_val=0.05, footprints=True):
    """All per-bar reductions of the reference for one set of close indices."""
    o = comp_bar_ohlcv(px, qty, idx)
    for k, name in enumerate(["open", "high", "low", "close", "volume", "vwap", "trades", "median"]):
        out[f"ref_{prefix}_ohlcv_{name}"] = o[k]
    d = comp_bar_directional_features(px, qty, idx, side)
    for k in range(14):
        out[f"ref_{prefix}_dir_{k}"] = d[k]
    nb = len(idx) - 1
    theta = np.full(nb, theta_val)
    if nb > 3:
        theta[3] = 0.0
    out[f"in_{prefix}_theta"] = theta
    t = comp_bar_trade_size_features(qty, theta, idx, 5.0)
    for k in range(4):
        out[f"ref_{prefix}_ts_{k}"] = t[k]
    if footprints:
        f = comp_bar_footprints(px, qty, idx, side, tick, o[2], o[1], 3.0)
        dts = [np.int32, np.float32, np.float32, np.int32, np.int32, np.bool_, np.bool_]
        for k in range(7):
            off, flat = csr(list(f[k]), dts[k])
            out[f"ref_{prefix}_fp_{k}"] = flat
        out[f"ref_{prefix}_fp_off"] = off
        for k in range(7, 13):
            out[f"ref_{prefix}_fp_{k}"] = np.asarray(f[k])


def case_stream(name, ts, px, qty, side, *, interval=60.0, tick_thr=100, vol_thr=5.0, dol_thr=1e5, tick=0.1,
                ret_window=60.0, half_life=60.0, tbm=True):
    out = {"in_ts": ts, "in_px": px, "in_qty": qty, "in_side": side,
           "in_params": np.array([interval, tick_thr, vol_thr, dol_thr, tick, ret_window, half_life])}
    clock, tidx = _time_bar_indexer(ts, interval)
    out["ref_time_clock"], out["ref_time_idx"] = clock, tidx
    out["ref_tick_idx"] = np.array(_tick_bar_indexer(ts, tick_thr), dtype=np.int64)
    out["ref_volume_idx"] = np.array(_volume_bar_indexer(qty, vol_thr), dtype=np.int64)
    out["ref_dollar_idx"] = np.array(_dollar_bar_indexer(px, qty, dol_thr), dtype=np.int64)
    ret = comp_lagged_returns(ts, px, ret_window, True)
    out["ref_lagret_log"] = ret
    out["ref_lagret_simple"] = comp_lagged_returns(ts, px, ret_window, False)
    sig = ewmst(ts, ret, half_life)
    out["ref_ewmst"] = sig
    sig_in = sig.copy()
    out["in_cusum_sigma"] = sig.copy()
    out["ref_cusum_idx"] = np.array(_cusum_bar_indexer(ts, px, sig_in, 5e-4, 2.0), dtype=np.int64)
    out["ref_cusum_sigma_filled"] = sig_in
    ref_bundle(ts, px, qty, side, tidx, "time", out, tick=tick)
    ref_bundle(ts, px, qty, side, out["ref_dollar_idx"], "dollar", out, tick=tick)
    ref_bundle(ts, px, qty, side, out["ref_volume_idx"], "volume", out, tick=tick)
    ref_bundle(ts, px, qty, side, out["ref_tick_idx"], "tick", out, tick=tick, footprints=False)
    if len(out["ref_cusum_idx"]) >= 2:
        ref_bundle(ts, px, qty, side, out["ref_cusum_idx"], "cusum", out, tick=tick, footprints=False)
    # a15: bar-level features on the dollar bars (returns with NaNs injected, buy/sell volumes from the directional tuple)
    dcl = out["ref_dollar_ohlcv_close"]
    bret = np.full(len(dcl), np.nan)
    bret[1:] = np.log(dcl[1:] / dcl[:-1])
    if len(bret) > 12:
        bret[[5, 11]] = np.nan
    out["in_bar_ret"] = bret
    for w, smp in [(5, True), (20, False)]:
        out[f"ref_rv_{w}_{int(smp)}"] = realized_vol(bret, w, smp)
    out["ref_ewms_10"] = ewms(bret, 10)
    vb, vs = out["ref_dollar_dir_2"].astype(np.float64), out["ref_dollar_dir_3"].astype(np.float64)
    if len(vb) > 9:
        vb[7] = np.nan
    out["in_vpin_vb"], out["in_vpin_vs"] = vb, vs
    out["ref_vpin_8"] = vpin(vb, vs, 8)
    out["ref_flow_acc_20_5"] = comp_flow_acceleration(out["ref_dollar_ohlcv_volume"].astype(np.float64), 20, 5)
    if tbm:
        # events: dollar-bar closes with a finite sigma, excluding those too close to the end
        ev = out["ref_dollar_idx"][1:]
        ev = ev[np.isfinite(sig[ev])]
        ev = ev[ts[ev] + int(interval * 5e9) <= ts[-1]]
        if len(ev) > 0:
            tg = sig[ev] * 1.0 + 1e-5
            out["in_tbm_events"], out["in_tbm_targets"] = ev, tg
            out["in_tbm_params"] = np.array([2.0, 2.0, interval * 5, 1.0, 0.0])
            lab, tch, rets, rat = triple_barrier(ts, px, ev, tg, (2.0, 2.0), interval * 5, 1.0, None, 0.0)
            out["ref_tbm_labels"], out["ref_tbm_touch"], out["ref_tbm_rets"], out["ref_tbm_ratios"] = lab, tch, rets, rat
            sd = np.where(np.arange(len(ev)) % 3 == 0, -1, 1).astype(np.int8)
            out["in_tbm_side"] = sd
            lab, tch, rets, rat = triple_barrier(ts, px, ev, tg, (1.0, np.inf), interval * 2, 0.0, sd, 1e-4)
            out["ref_tbm_meta_labels"], out["ref_tbm_meta_touch"] = lab, tch
            out["ref_tbm_meta_rets"], out["ref_tbm_meta_ratios"] = rets, rat
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    nb = {k: len(out[f"ref_{k}_idx"]) - 1 for k in ["time", "tick", "volume", "dollar", "cusum"]}
    print(f"{name}: n={len(ts)} bars={nb} events={len(out.get('in_tbm_events', []))}")


def weights_cases():
    """SURVEY 8f-1 sample weights: reference outputs on the synth_20k TBM events, an int16 wrap-around case and a series
    with a zero close (infinite log return)."""
    from finmlkit.label.weights import average_uniqueness, return_attribution, time_decay, class_balance_weights
    g = dict(np.load(os.path.join(HERE, "synth_20k.npz")))
    out = {}
    ts, px = g["in_ts"], g["in_px"]
    ev, tc = g["in_tbm_events"], g["ref_tbm_touch"]
    out["a_ts"], out["a_px"], out["a_ev"], out["a_touch"] = ts, px, ev, tc
    w, c = average_uniqueness(ts, ev, tc)
    out["a_ref_avg_u"], out["a_ref_conc"] = w, c
    out["a_ref_ra"] = return_attribution(ev, tc, px, c, False)
    out["a_ref_ra_norm"] = return_attribution(ev, tc, px, c, True)
    out["a_ref_decay_05"] = time_decay(w, 0.5)
    out["a_ref_decay_m03"] = time_decay(w, -0.3)
    lab = g["ref_tbm_labels"]
    cb = class_balance_weights(lab, w)
    out["a_labels"] = lab
    for k in range(4):
        out[f"a_ref_cb_{k}"] = cb[k]
    # int16 wrap: 70 000 identical labels over ticks 10..20 plus a few ordinary ones; a zero close inside a label
    n = 400
    rng = np.random.default_rng(3)
    ts2 = np.arange(n, dtype=np.int64)
    px2 = np.round(100 + np.cumsum(rng.normal(0, 0.05, n)), 2)
    px2[300] = 0.0
    ev2 = np.concatenate([np.full(70000, 10), np.array([0, 5, 15, 100, 250, 290, 310, 399])]).astype(np.int64)
    tc2 = np.concatenate([np.full(70000, 20), np.array([30, 5, 18, 260, 299, 305, 398, 399])]).astype(np.int64)
    w2, c2 = average_uniqueness(ts2, ev2, tc2)
    out["b_ts"], out["b_px"], out["b_ev"], out["b_touch"] = ts2, px2, ev2, tc2
    out["b_ref_avg_u"], out["b_ref_conc"] = w2, c2
    out["b_ref_ra"] = return_attribution(ev2, tc2, px2, c2, False)
    # 65 536 identical labels: the wrapped concurrency is exactly 0 inside the labels -> inf uniqueness
    ev3 = np.concatenate([np.full(65536, 10), np.array([12, 50])]).astype(np.int64)
    tc3 = np.concatenate([np.full(65536, 20), np.array([60, 70])]).astype(np.int64)
    w3, c3 = average_uniqueness(ts2, ev3, tc3)
    out["c_ev"], out["c_touch"], out["c_ref_avg_u"], out["c_ref_conc"] = ev3, tc3, w3, c3
    out["c_ref_ra"] = return_attribution(ev3, tc3, px2, c3, False)
    np.savez_compressed(os.path.join(HERE, "weights.npz"), **out)
    print("weights: events", len(ev), "max concurrency", int(c.max()), "wrap conc", int(c2[15]), "zero-wrap", int(c3[15]), w3[:1])


def ingest_cases():
    """cusum_filter (constant and per-tick thresholds, NaN threshold stretch), tick rule, split-trade merging."""
    from finmlkit.sampling.filters import cusum_filter
    from finmlkit.bar.utils import comp_trade_side_vector, merge_split_trades
    out = {}
    ts, px, qty, side = synth_trades(50000, seed=21)
    out["f_px"] = px
    out["f_thr_const"] = np.array([2e-4])
    out["f_ref_const"] = cusum_filter(px, out["f_thr_const"])
    rng = np.random.default_rng(5)
    thr = np.abs(rng.normal(3e-4, 1e-4, len(px)))
    thr[1000:1200] = np.nan
    out["f_thr_arr"] = thr
    out["f_ref_arr"] = cusum_filter(px, thr)
    # tick rule: flat stretches, sub-epsilon moves
    p2 = px.copy()
    p2[100:140] = p2[100]
    p2[200] = p2[199] + 5e-13
    out["s_px"] = p2
    out["s_ref"] = comp_trade_side_vector(p2)
    # merging: duplicate timestamps with split fills; prices within / beyond 1e-8 of the head; both sides
    n = 20000
    ts3 = 1_700_000_000_000_000_000 + np.cumsum(rng.integers(0, 3, n) * (rng.random(n) < 0.4)).astype(np.int64) * 1_000_000
    px3 = np.round(100 + rng.integers(0, 3, n) * 0.5, 1) + rng.choice([0.0, 4e-9, 9e-9, 2e-8], n)
    ibm = rng.random(n) < 0.5
    order = np.lexsort((ibm, px3, ts3))
    ts3, px3, ibm = ts3[order], px3[order], ibm[order]
    am3 = np.round(rng.lognormal(-3, 1, n), 3).astype(np.float32)
    out["m_ts"], out["m_px"], out["m_am"], out["m_ibm"] = ts3, px3, am3, ibm
    r = merge_split_trades(ts3, px3, am3, ibm)
    for k in range(4):
        out[f"m_ref_{k}"] = np.asarray(r[k])
    r = merge_split_trades(ts3, px3, am3, None)
    for k in range(3):
        out[f"m_ref_noside_{k}"] = np.asarray(r[k])
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), **out)
    print("ingest: events", len(out["f_ref_const"]), len(out["f_ref_arr"]), "merged", len(out["m_ref_0"]), "of", n, "/", len(out["m_ref_noside_0"]))


def volprofile_cases():
    """volume_profile_rolling of the reference on dollar-bar footprints of a 150k-tick stream (tick 0.1)."""
    from numba.typed import List as NL
    from finmlkit.feature.core.volume import volume_profile_rolling
    ts, px, qty, side = synth_trades(150000, seed=8)
    idx = np.array(_dollar_bar_indexer(px, qty, 3e4), dtype=np.int64)
    o = comp_bar_ohlcv(px, qty, idx)
    f = comp_bar_footprints(px, qty, idx, side, 0.1, o[2], o[1], 3.0)
    off, lv = csr(list(f[0]), np.int32)
    _, bv = csr(list(f[1]), np.float32)
    _, sv = csr(list(f[2]), np.float32)
    bts = ts[idx[1:]]
    out = {"ts": bts, "high": o[1], "low": o[2], "off": off, "levels": lv, "buy": bv, "sell": sv}
    cases = [(60.0, 27), (300.0, 0), (20.0, 5), (600.0, 27), (5.0, 27), (120.0, 3), (60.0, 200)]
    out["cases"] = np.array(cases)
    for k, (w, nbins) in enumerate(cases):
        r = volume_profile_rolling(bts, o[1], o[2], f[0], f[1], f[2], w, int(nbins) if nbins else None, 0.1, 68.34)
        for q in range(4):
            out[f"ref_{k}_{q}"] = r[q]
    # single-level windows (constant price) and a one-bar series
    n = 64
    ts2 = np.arange(n, dtype=np.int64) * 1_000_000_000
    pl, bl, sl = NL(), NL(), NL()
    for i in range(n):
        pl.append(np.array([1000], np.int32)); bl.append(np.array([1.0 + i], np.float32)); sl.append(np.array([0.5], np.float32))
    r = volume_profile_rolling(ts2, np.full(n, 100.0), np.full(n, 100.0), pl, bl, sl, 5.0, 27, 0.1, 68.34)
    for q in range(4):
        out[f"flat_ref_{q}"] = r[q]
    np.savez_compressed(os.path.join(HERE, "volprofile.npz"), **out)
    print("volprofile: bars", len(bts), "levels", len(lv), "flat poc", r[0][:8])


def main():
    if "--only-volprofile" in sys.argv:
        volprofile_cases()
        return
    if "--only-weights" in sys.argv:
        weights_cases()
        return
    if "--only-ingest" in sys.argv:
        ingest_cases()
        return
    # 1. plain synthetic stream (same generator the bench uses)
    ts, px, qty, side = synth_trades(20000, seed=42)
    case_stream("synth_20k", ts, px, qty, side)

    # 2. adversarial: quantised equal-ish sizes -> exact-tie volume thresholds; hour-long gaps -> empty time bars;
    #    first tick 100 ns before a minute boundary (H1); side == 0 ticks; duplicate timestamps
    rng = np.random.default_rng(7)
    n = 6000
    gaps = rng.integers(1, 40, n).astype(np.int64) * 1_000_000
    gaps[rng.random(n) < 0.3] = 0                      # duplicate timestamps
    gaps[[1500, 3000]] = 3_600_000_000_000 + 17        # hour-long gaps
    ts = 1_600_000_019_999_999_900 + np.cumsum(gaps)
    ts[0] = 1_600_000_019_999_999_900
    px = np.round(100.0 + np.cumsum(rng.choice([-0.5, 0.0, 0.0, 0.5], n)), 1)
    qty = rng.choice([0.001, 0.002, 0.005, 0.01, 0.1, 0.25], n)
    side = rng.choice(np.array([-1, 1, 1, -1, 0], dtype=np.int8), n)
    case_stream("adversarial_6k", ts, px, qty, side, interval=60.0, tick_thr=7, vol_thr=1.0, dol_thr=250.0, tick=0.5,
                ret_window=5.0, half_life=30.0)

    # 3. giant trades: single ticks worth several dollar/volume thresholds, threshold-1 tick bars, tiny thresholds
    rng = np.random.default_rng(11)
    n = 4000
    ts = 1_700_000_000_000_000_000 + np.cumsum(rng.integers(0, 3, n)).astype(np.int64) * 250_000_000
    px = np.round(50.0 * np.exp(np.cumsum(rng.normal(0, 1e-3, n))), 2)
    qty = np.round(rng.lognormal(0, 2.0, n), 3) + 0.001
    side = rng.choice(np.array([-1, 1], dtype=np.int8), n)
    case_stream("giant_4k", ts, px, qty, side, interval=1.0, tick_thr=1, vol_thr=3.0, dol_thr=150.0, tick=0.01,
                ret_window=2.0, half_life=10.0)

    # 4. sub-second clock (H12: numba arange evaluates int64(start + i*step) in float64)
    ts, px, qty, side = synth_trades(3000, seed=5)
    for iv in [0.001, 0.0005, 0.25]:
        clock, idx = _time_bar_indexer(ts, iv)
        np.savez_compressed(os.path.join(HERE, f"clock_{iv}.npz"), in_ts=ts, in_interval=np.array([iv]),
                            ref_time_clock=clock, ref_time_idx=idx)

    weights_cases()
    ingest_cases()
    volprofile_cases()
    crosscheck()


def crosscheck():
    """C oracle vs reference on 1M ticks (not stored)."""
    import oracle
    ts, px, qty, side = synth_trades(1_000_000, seed=1)
    ok = True

    def chk(name, a, b, exact=True, atol=0):
        nonlocal ok
        a, b = np.asarray(a), np.asarray(b)
        good = a.shape == b.shape and (np.array_equal(a, b, equal_nan=True) if exact else np.allclose(a, b, rtol=1e-12, atol=atol, equal_nan=True))
        ok &= bool(good)
        print(f"  {name:28s} {'OK' if good else 'MISMATCH'}")

    c, i = _time_bar_indexer(ts, 60.0); oc, oi = oracle.time_bar_indexer(ts, 60.0)
    chk("time clock", c, oc); chk("time idx", i, oi)
    chk("tick", np.array(_tick_bar_indexer(ts, 1000)), oracle.tick_bar_indexer(ts, 1000))
    chk("volume", np.array(_volume_bar_indexer(qty, 5.0)), oracle.volume_bar_indexer(qty, 5.0))
    di = np.array(_dollar_bar_indexer(px, qty, 1e5), dtype=np.int64)
    chk("dollar", di, oracle.dollar_bar_indexer(px, qty, 1e5))
    r = comp_lagged_returns(ts, px, 60.0, True); chk("lagret", r, oracle.comp_lagged_returns(ts, px, 60.0, True))
    s = ewmst(ts, r, 60.0); chk("ewmst", s, oracle.ewmst(ts, r, 60.0))
    s1, s2 = s.copy(), s.copy()
    chk("cusum", np.array(_cusum_bar_indexer(ts, px, s1, 5e-4, 2.0)), oracle.cusum_bar_indexer(ts, px, s2, 5e-4, 2.0))
    for nm, idx in [("dollar", di), ("time", i)]:
        a, b = comp_bar_ohlcv(px, qty, idx), oracle.comp_bar_ohlcv(px, qty, idx)
        for k in range(8):
            chk(f"ohlcv[{nm}][{k}]", a[k], b[k])
        a, b = comp_bar_directional_features(px, qty, idx, side), oracle.comp_bar_directional_features(px, qty, idx, side)
        for k in range(14):
            chk(f"dir[{nm}][{k}]", a[k], b[k])
        th = np.full(len(idx) - 1, 0.05)
        a, b = comp_bar_trade_size_features(qty, th, idx, 5.0), oracle.comp_bar_trade_size_features(qty, th, idx, 5.0)
        for k in range(4):
            chk(f"tsize[{nm}][{k}]", a[k], b[k])
        o = comp_bar_ohlcv(px, qty, idx)
        a = comp_bar_footprints(px, qty, idx, side, 0.1, o[2], o[1], 3.0)
        b = oracle.comp_bar_footprints_csr(px, qty, idx, side, 0.1, o[2], o[1], 3.0)
        dts = [np.int32, np.float32, np.float32, np.int32, np.int32, np.bool_, np.bool_]
        for k in range(7):
            chk(f"fp[{nm}][{k}]", csr(list(a[k]), dts[k])[1], b[1 + k])
        for k in range(7, 13):
            chk(f"fp[{nm}][{k}]", a[k], b[1 + k], exact=(k != 11), atol=(1e-3 if k == 11 else 0))
    ev = di[1:-50]; ev = ev[np.isfinite(s[ev])]
    tg = s[ev] + 1e-5
    a = triple_barrier(ts, px, ev, tg, (2.0, 2.0), 300.0, 1.0, None, 0.0)
    b = oracle.triple_barrier(ts, px, ev, tg, (2.0, 2.0), 300.0, 1.0, None, 0.0)
    for k in range(4):
        chk(f"tbm[{k}]", a[k], b[k])
    g = dict(np.load(os.path.join(HERE, "synth_20k.npz")))
    chk("a15 realized_vol", g["ref_rv_5_1"], oracle.realized_vol(g["in_bar_ret"], 5, True))
    chk("a15 realized_vol pop", g["ref_rv_20_0"], oracle.realized_vol(g["in_bar_ret"], 20, False))
    chk("a15 ewms", g["ref_ewms_10"], oracle.ewms(g["in_bar_ret"], 10))
    chk("a15 vpin", g["ref_vpin_8"], oracle.vpin(g["in_vpin_vb"], g["in_vpin_vs"], 8))
    chk("a15 flow acc", g["ref_flow_acc_20_5"], oracle.comp_flow_acceleration(g["ref_dollar_ohlcv_volume"].astype(np.float64), 20, 5))
    print("CROSSCHECK", "PASSED" if ok else "FAILED")


if __name__ == "__main__":
    main()
