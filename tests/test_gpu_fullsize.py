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


def test_dollar_1e9_properties_and_prefix_parity(full, ctx):
    from finmlkit_b200 import core
    tr, ix, stats = full
    cts, cidx = ix.download()
    n = N_FULL
    assert cidx[0] == 0 and np.all(np.diff(cidx) > 0) and cidx[-1] < n          # sorted, strictly increasing
    # prefix of the same stream on the host -> oracle (serial C restatement of logic.py:118-149)
    npre = min(N_PREFIX, n)
    ts = np.empty(n, np.int64); px = np.empty(n); qty = np.empty(n)
    tr.download(out=(ts, px, qty, None))
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
    for ch in ("32", "96"):
        monkeypatch.setenv("FMK_CUSUM_CH", ch)
        for name in STREAM_CASES:
            g = load_case(name)
            tr = core.DeviceTrades.upload(g["in_ts"], g["in_px"], g["in_qty"], g["in_side"], ctx=ctx)
            sig = core.DeviceBuf.upload(ctx, g["in_cusum_sigma"])
            idx = core.cusum_bar_index(tr, sig, 5e-4, 2.0).download()[1]
            assert np.array_equal(idx, g["ref_cusum_idx"]), f"{name} CH={ch}"
            st = ctx.index_stats()
            assert st["tasks"] > 30
    monkeypatch.delenv("FMK_CUSUM_CH")


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
    for floor, mult in ((5e-4, 2.0), (1e-5, 0.05)):
        sb = core.DeviceBuf.upload(ctx, sigma)
        idx = core.cusum_bar_index(tr, sb, floor, mult).download()[1]
        ref = oracle.cusum_bar_indexer(ts, px, sigma.copy(), floor, mult)
        assert np.array_equal(idx, ref), (floor, mult, len(idx), len(ref))
