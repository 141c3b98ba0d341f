"""GPU: BASELINE-size streams.  Direct oracle comparison on a 1e8-tick stream, and at 1e9 ticks size-independent
properties plus CAUSALITY: the indices a 1e9-tick run produces below tick 1e8 must equal the oracle's indices on the
1e8 prefix alone (every indexer here is causal), which ties the full-size run to the oracle bit for bit."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

N_FULL = int(float(os.environ.get("FMK_FULLSIZE_TICKS", "1e9")))
N_PREFIX = 100_000_000
T = 1e6


@pytest.fixture(scope="module")
def full(ctx):
    from finmlkit_b200 import core
    tr = core.DeviceTrades.synth(N_FULL, seed=42, ctx=ctx)
    ix = core.dollar_bar_index(tr, T)
    stats = ctx.index_stats()
    return tr, ix, stats


@pytest.fixture(scope="module")
def host(full):
    """the full stream on the host (25 B/tick), downloaded once for every test of this module"""
    tr = full[0]
    n = N_FULL
    ts = np.empty(n, np.int64); px = np.empty(n); qty = np.empty(n); side = np.empty(n, np.int8)
    tr.download(out=(ts, px, qty, side))
    return ts, px, qty, side


def test_dollar_1e9_properties_and_prefix_parity(full, host, ctx):
    from finmlkit_b200 import core
    tr, ix, stats = full
    cts, cidx = ix.download()
    n = N_FULL
    assert cidx[0] == 0 and np.all(np.diff(cidx) > 0) and cidx[-1] < n          # sorted, strictly increasing
    # prefix of the same stream on the host -> oracle (serial C restatement of logic.py:118-149)
    npre = min(N_PREFIX, n)
    ts, px, qty, _ = host
    ref = oracle.dollar_bar_indexer(px[:npre], qty[:npre], T)
    got = cidx[cidx < npre]
    assert np.array_equal(got, ref), "causality/prefix parity with the oracle failed"
    assert np.array_equal(cts, ts[cidx])
    # exact-arithmetic count: every emission removes T from the running dollar sum
    total = float(np.sum(px * qty))
    assert abs((len(cidx) - 1) - int(total // T)) <= 1
    # the fast path must carry the headline workload (no serial repairs on this stream)
    assert stats["serial_repairs"] <= 2, stats
    # OHLCV checksum of checksums on the device-resident result
    o = core.bar_ohlcv(tr, ix)
    assert int(o[6].sum()) == int(cidx[-1] - cidx[0])                             # trades partition the covered ticks
    covered = slice(int(cidx[0]) + 1, int(cidx[-1]) + 1)
    assert np.isclose(float(o[4].astype(np.float64).sum()), float(qty[covered].sum()), rtol=1e-6)
    assert np.all(o[1] >= o[2]) and np.all(o[1] >= o[0]) and np.all(o[2] <= o[3])
    # idempotence / determinism: a second build is bit-identical
    c2 = core.dollar_bar_index(tr, T).download()[1]
    assert np.array_equal(c2, cidx)
    # full oracle on the complete stream when the host has the memory for it (16 B/tick already resident here)
    if n <= 1_000_000_000:
        ref_full = oracle.dollar_bar_indexer(px, qty, T)
        assert np.array_equal(cidx, ref_full), "bit-exact index check vs the oracle at full size failed"
        oo = oracle.comp_bar_ohlcv(px[:npre], qty[:npre], ref)
        k = len(ref) - 1
        for col in (0, 1, 2, 3, 6, 7):
            assert np.array_equal(o[col][:k], oo[col]), f"ohlcv column {col}"
        assert np.allclose(o[5][:k], oo[5], rtol=1e-9, atol=0)


def test_time_and_volume_bars_1e8(ctx):
    from finmlkit_b200 import core
    n = min(N_PREFIX, N_FULL)
    tr = core.DeviceTrades.synth(n, seed=43, ctx=ctx)
    ts, px, qty, side = tr.download()
    clock, tidx = core.time_bar_index(tr, 60.0).download()
    rc, ri = oracle.time_bar_indexer(ts, 60.0)
    assert np.array_equal(clock, rc) and np.array_equal(tidx, ri)
    vidx = core.volume_bar_index(tr, 50.0).download()[1]
    assert np.array_equal(vidx, oracle.volume_bar_indexer(qty, 50.0))
    tk = core.tick_bar_index(tr, 1000).download()[1]
    assert np.array_equal(tk, oracle.tick_bar_indexer(ts, 1000))


def test_cusum_fixpoint_small_chunks(ctx, monkeypatch):
    """CUSUM chunk chain with 32/96-tick chunks: hundreds of chunks whose state does not coalesce within one chunk, so the
    parallel fix-point needs many rounds -- the result must still be the sequential trajectory (golden reference)."""
    from finmlkit_b200 import core
    from helpers import STREAM_CASES, load_case
    # every combination of: chunk size, lane-per-chunk kernel only / warp walkers only / default hand-over, walker reach,
    # integer-compare fast chain on / off
    for ch, walk_below, walk_max, no_fast in (("32", None, None, None), ("96", None, None, None), ("32", "0", None, None),
                                              ("96", "1000000000", "1", None), ("32", "1000000000", "7", None),
                                              ("96", "1000000000", "2", "1"), ("32", "0", None, "1")):
        monkeypatch.setenv("FMK_CUSUM_CH", ch)
        for var, val in (("FMK_CUSUM_WALK_BELOW", walk_below), ("FMK_CUSUM_WALK_MAX", walk_max), ("FMK_CUSUM_NO_FAST", no_fast)):
            if val is None:
                monkeypatch.delenv(var, raising=False)
            else:
                monkeypatch.setenv(var, val)
        for name in STREAM_CASES:
            g = load_case(name)
            tr = core.DeviceTrades.upload(g["in_ts"], g["in_px"], g["in_qty"], g["in_side"], ctx=ctx)
            sig = core.DeviceBuf.upload(ctx, g["in_cusum_sigma"])
            idx = core.cusum_bar_index(tr, sig, 5e-4, 2.0).download()[1]
            assert np.array_equal(idx, g["ref_cusum_idx"]), f"{name} CH={ch} walk_below={walk_below} walk_max={walk_max} no_fast={no_fast}"
            st = ctx.index_stats()
            assert st["tasks"] > 30
    for var in ("FMK_CUSUM_CH", "FMK_CUSUM_WALK_BELOW", "FMK_CUSUM_WALK_MAX", "FMK_CUSUM_NO_FAST"):
        monkeypatch.delenv(var, raising=False)


def test_cusum_1e7_vs_oracle(ctx):
    """1-hour sigma pipeline + CUSUM bars on 1e7 ticks (2441 chunks, multi-round repair) against the oracle given the same
    sigma; a low threshold variant closes bars every few hundred ticks."""
    import ctypes as C
    from finmlkit_b200 import core
    n = 10_000_000
    tr = core.DeviceTrades.synth(n, seed=44, ctx=ctx)
    ts, px, qty, side = tr.download()
    L = ctx._L
    r, s = C.c_void_p(), C.c_void_p()
    ctx.check(L.fmk_lagged_returns_dev(ctx.h, tr.h, 3600.0, 1, C.byref(r)))
    ctx.check(L.fmk_ewmst_dev(ctx.h, tr.h, r, 3600.0, 1e-12, C.byref(s)))
    L.fmk_buf_free(ctx.h, r)
    sigma = core.DeviceBuf(ctx, s).download(np.float64, n)
    for floor, mult in ((5e-4, 2.0), (1e-5, 0.05), (0.0, 1.0), (-1.0, 1.5)):     # floor < 0: the generic (double-compare) chain
        sb = core.DeviceBuf.upload(ctx, sigma)
        idx = core.cusum_bar_index(tr, sb, floor, mult).download()[1]
        ref = oracle.cusum_bar_indexer(ts, px, sigma.copy(), floor, mult)
        assert np.array_equal(idx, ref), (floor, mult, len(idx), len(ref))


def test_config3_full_size_volume_directional_footprints(full, host, ctx):
    """BASELINE configs[2] at full size inside the driver-run suite: volume bars + directional + footprint CSR on the whole
    stream; every bar that closes inside the first 1e8 ticks must equal the oracle run on that prefix alone (the indexer and
    the per-bar reductions are causal), and size-independent properties hold on the rest (int32 next[] tables, hundreds of
    millions of CSR rows and their int64 offsets are exactly what breaks at size)."""
    from finmlkit_b200 import core
    from helpers import check_directional, check_footprint_csr, check_ohlcv
    tr = full[0]
    ts, px, qty, side = host
    n = N_FULL
    npre = min(N_PREFIX, n)
    vix = core.volume_bar_index(tr, 50.0)
    cts, cidx = vix.download()
    ref = oracle.volume_bar_indexer(qty[:npre], 50.0)
    k = len(ref) - 1
    assert np.array_equal(cidx[:k + 1], ref) and (k + 1 == len(cidx) or cidx[k + 1] >= npre), "volume prefix parity"
    assert cidx[0] == 0 and np.all(np.diff(cidx) > 0) and cidx[-1] < n
    assert np.array_equal(cts, ts[cidx])
    fr = core.bar_features_device(tr, vix, core.F_OHLCV | core.F_MEDIAN | core.F_DIRECTIONAL | core.F_FOOTPRINT,
                                  price_tick_size=0.1, imbalance_factor=3.0)
    c = fr.download()
    nb = len(cidx) - 1
    # ---- prefix parity against the oracle on the first 1e8 ticks ----
    oo = oracle.comp_bar_ohlcv(px[:npre], qty[:npre], ref)
    got_o = [c[x][:k] for x in ("open", "high", "low", "close", "volume", "vwap", "trades", "median_trade_size")]
    check_ohlcv(got_o, oo, "cfg3 prefix")
    dn = ["ticks_buy", "ticks_sell", "volume_buy", "volume_sell", "dollars_buy", "dollars_sell", "mean_spread", "max_spread",
          "cum_ticks_min", "cum_ticks_max", "cum_volume_min", "cum_volume_max", "cum_dollars_min", "cum_dollars_max"]
    check_directional([c[x][:k] for x in dn], oracle.comp_bar_directional_features(px[:npre], qty[:npre], ref, side[:npre]), "cfg3 prefix")
    fo = oracle.comp_bar_footprints_csr(px[:npre], qty[:npre], ref, side[:npre], 0.1, oo[2], oo[1], 3.0)
    off = c["fp_level_offsets"]
    nl = int(off[k])
    got_fp = (off[:k + 1], c["fp_price_levels"][:nl], c["fp_buy_vol"][:nl], c["fp_sell_vol"][:nl], c["fp_buy_ticks"][:nl],
              c["fp_sell_ticks"][:nl], c["fp_buy_imb"][:nl], c["fp_sell_imb"][:nl], c["fp_buy_imb_sum"][:k], c["fp_sell_imb_sum"][:k],
              c["fp_cot"][:k], c["fp_run_signed"][:k], c["fp_vp_skew"][:k], c["fp_vp_gini"][:k])
    check_footprint_csr(got_fp, fo[0], list(fo[1:]), float(np.max(np.abs(fo[1]))), "cfg3 prefix")
    # ---- size-independent properties on the whole stream ----
    assert int(c["trades"].sum()) == int(cidx[-1] - cidx[0])
    assert np.array_equal(c["ticks_buy"] + c["ticks_sell"], c["trades"])           # the synthetic side column is +-1
    assert off[0] == 0 and np.all(np.diff(off) >= 1) and off[-1] == fr.n_levels
    lv_per_bar = np.rint(c["high"] / 0.1).astype(np.int64) - np.rint(c["low"] / 0.1).astype(np.int64) + 1
    assert np.array_equal(np.diff(off), lv_per_bar)
    lvl_ticks = np.add.reduceat(c["fp_buy_ticks"].astype(np.int64) + c["fp_sell_ticks"], off[:-1])
    assert np.array_equal(lvl_ticks, c["trades"])                                   # every tick landed on exactly one level
    first_lv = c["fp_price_levels"][off[:-1]]
    assert np.array_equal(first_lv, np.rint(c["low"] / 0.1).astype(np.int64))
    assert nb == len(c["open"]) and fr.n_levels > nb


def test_config4_full_size_sigma_cusum_tbm_weights(full, host, ctx):
    """BASELINE configs[3] (oracle-pinned half) at full size: sigma = ewmst(lagged log returns) -> CUSUM bars -> triple barrier
    -> sample weights on the whole stream; everything that depends only on the first 1e8 ticks must equal the oracle run on
    that prefix (sigma within 1e-9, CUSUM indices and TBM labels / touch indices bit-exact, weights within 1e-9)."""
    from finmlkit_b200 import core
    tr = full[0]
    ts, px, qty, side = host
    n = N_FULL
    npre = min(N_PREFIX, n)
    r = core.lagged_returns_dev(tr, 3600.0, True)
    sig = core.ewmst_dev(tr, r, 3600.0)
    sig_pre = sig.download(np.float64, n)[:npre].copy()
    rr = oracle.comp_lagged_returns(ts[:npre], px[:npre], 3600.0, True)
    assert_f64_(r.download(np.float64, n)[:npre], rr, "lagged returns prefix", 1e-9, 1e-15)
    del r
    assert_f64_(sig_pre, oracle.ewmst(ts[:npre], rr, 3600.0), "sigma prefix", 1e-9, 1e-18)
    cix = core.cusum_bar_index(tr, sig, 5e-4, 2.0)                   # forward-fills the device sigma in place, like the reference
    stats = ctx.index_stats()
    cts, cidx = cix.download()
    cref = oracle.cusum_bar_indexer(ts[:npre], px[:npre], sig_pre.copy(), 5e-4, 2.0)   # same sigma in: the indexer alone is compared
    kk = len(cref)
    assert np.array_equal(cidx[:kk], cref) and (kk == len(cidx) or cidx[kk] >= npre), "cusum prefix parity"
    assert np.all(np.diff(cidx) > 0) and cidx[-1] < n and np.array_equal(cts, ts[cidx])
    assert stats["tasks"] > 1000
    # events = bar closes with a finite target whose vertical barrier fits the stream (label/kit.py:262-269)
    ev = cidx[1:]
    tg = sig.gather(ev)
    keep = np.isfinite(tg) & (ts[ev] + 3600 * 10**9 <= ts[-1])
    ev, tg = ev[keep], tg[keep]
    lab = core.triple_barrier_dev(tr, ev, tg, (2.0, 2.0), 3600.0, 1.0, None, 0.0)
    assert np.all(lab[1] >= ev) and np.all(lab[1] < n) and set(np.unique(lab[0])) <= {-1, 1}
    # prefix events: the whole 1-hour path lies inside the prefix
    sel = ts[ev] + 3600 * 10**9 < ts[npre - 1]
    evp, tgp = ev[sel], tg[sel]
    assert len(evp) > 100
    lr = oracle.triple_barrier(ts[:npre], px[:npre], evp, tgp, (2.0, 2.0), 3600.0, 1.0, None, 0.0)
    m = len(evp)
    assert np.array_equal(lab[0][:m], lr[0]) and np.array_equal(lab[1][:m], lr[1]), "tbm prefix labels / touch indices"
    assert_f64_(lab[2][:m], lr[2], "tbm returns", 1e-9, 1e-15)
    assert_f64_(lab[3][:m], lr[3], "tbm ratios", 1e-9, 1e-15)
    # sample weights on the full event list; the prefix events' concurrency only involves prefix events (paths <= 1 h)
    u, ra = core.sample_weights_dev(tr, ev, lab[1])
    assert np.all(u > 0) and np.all(u <= 1.0) and np.all(np.isfinite(ra))
    sel2 = ts[ev] + 2 * 3600 * 10**9 < ts[npre - 1]                  # labels that could overlap them also end inside the prefix
    m2 = int(sel2.sum())
    ow, oc = oracle.average_uniqueness(ts[:npre], ev[:m], lab[1][:m])
    assert_f64_(u[:m2], ow[:m2], "avg uniqueness prefix", 1e-9, 1e-15)
    ora = oracle.return_attribution(ev[:m], lab[1][:m], px[:npre], oc, False)
    assert_f64_(ra[:m2], ora[:m2], "return attribution prefix", 1e-9, 1e-13)


def test_time_bars_full_size_north_star(full, host, ctx):
    """the north-star workload (1-minute time bars + OHLCV incl. median) on the whole stream: clock and indices against the
    oracle in full (the time-bar indexer is one vectorised search), OHLCV against the oracle on the 1e8 prefix."""
    from finmlkit_b200 import core
    from helpers import check_ohlcv
    tr = full[0]
    ts, px, qty, side = host
    n = N_FULL
    tix = core.time_bar_index(tr, 60.0)
    clock, tidx = tix.download()
    rc, ri = oracle.time_bar_indexer(ts, 60.0)
    assert np.array_equal(clock, rc) and np.array_equal(tidx, ri)
    o = core.bar_ohlcv(tr, tix)
    npre = min(N_PREFIX, n)
    kb = int(np.searchsorted(tidx, npre - 1, "left"))               # bars whose close index lies inside the prefix
    oo = oracle.comp_bar_ohlcv(px[:npre], qty[:npre], tidx[:kb])
    check_ohlcv([x[:kb - 1] for x in o], oo, "time bars prefix")
    assert int(o[6].sum()) == int(tidx[-1] - tidx[0])


def assert_f64_(a, b, what, rtol, atol):
    from helpers import assert_f64
    assert_f64(a, b, what, rtol=rtol, atol=atol)
