"""Edge cases of the widened rows on the GPU: empty / single-element / degenerate inputs, against the oracle."""
import numpy as np
import pytest

from helpers import assert_exact

pytestmark = pytest.mark.gpu


def test_weights_edges(ctx):
    import oracle
    from finmlkit_b200.label import weights as W
    ts = np.arange(8, dtype=np.int64)
    empty = np.zeros(0, np.int64)
    w, c = W.average_uniqueness(ts, empty, empty, ctx=ctx)
    assert len(w) == 0 and c.dtype == np.int16 and not c.any()
    ra = W.return_attribution(empty, empty, np.ones(8), np.zeros(8, np.int16), False, ctx=ctx)
    assert len(ra) == 0 and ra.dtype == np.float64
    # a single label covering the whole (one-tick) series, and one covering everything
    w, c = W.average_uniqueness(ts[:1], np.array([0]), np.array([0]), ctx=ctx)
    assert_exact(w, np.array([1.0]), "single tick")
    px = np.array([100., 101., 99., 0., 5., 5., 6., 7.])        # a zero close inside the label: log -> -inf, next one skipped
    ev, tc = np.array([0, 2]), np.array([7, 5])
    ow, oc = oracle.average_uniqueness(ts, ev, tc)
    w, c = W.average_uniqueness(ts, ev, tc, ctx=ctx)
    assert_exact(c, oc, "conc")
    np.testing.assert_allclose(w, ow, rtol=1e-12)
    ora = oracle.return_attribution(ev, tc, px, oc, False)
    ra = W.return_attribution(ev, tc, px, c, False, ctx=ctx)
    assert np.array_equal(np.isinf(ra), np.isinf(ora)) and np.all(np.isinf(ra))


def test_filter_and_ingest_edges(ctx):
    import oracle
    from finmlkit_b200.bar.utils import comp_trade_side_vector, merge_split_trades
    from finmlkit_b200.sampling.filters import cusum_filter
    assert len(cusum_filter(np.array([100.0, 100.0]), np.array([0.1]), ctx=ctx)) == 0
    assert_exact(cusum_filter(np.array([100.0, 150.0]), np.array([0.1]), ctx=ctx), np.array([1], np.int64), "two ticks")
    flat = np.full(5000, 42.0)
    assert len(cusum_filter(flat, np.array([1e-9]), ctx=ctx)) == 0
    x = np.array([1.0, 2.0, 0.0, 3.0, np.nan, 4.0, 5.0, 2.5, 2.5, 10.0])         # zero / NaN prices: inf / NaN log returns
    assert_exact(cusum_filter(x, np.array([0.5]), ctx=ctx), oracle.cusum_filter(x, np.array([0.5])), "non-finite returns")
    assert_exact(comp_trade_side_vector(np.array([7.0]), ctx=ctx), np.array([0], np.int8), "one price")
    assert_exact(comp_trade_side_vector(np.array([1.0, 1.0, 1.0]), ctx=ctx), np.zeros(3, np.int8), "flat")
    one = merge_split_trades(np.array([5], np.int64), np.array([1.5]), np.array([2.0], np.float32), np.array([True]), ctx=ctx)
    assert_exact(one[0], np.array([5], np.int64), "ts")
    assert_exact(one[3], np.array([-1], np.int8), "side")
    n = 3000                                           # one timestamp run, one price, one side -> a single merged trade
    r = merge_split_trades(np.full(n, 9, np.int64), np.full(n, 2.0), np.full(n, 0.1, np.float32), np.zeros(n, bool), ctx=ctx)
    o = oracle.merge_split_trades(np.full(n, 9, np.int64), np.full(n, 2.0), np.full(n, 0.1, np.float32), np.zeros(n, bool))
    for k in range(4):
        assert_exact(r[k], o[k], f"single run [{k}]")
    assert len(r[0]) == 1


def test_volume_profile_edges(ctx):
    import oracle
    from finmlkit_b200 import core
    # one bar: no full window -> zeros, like the reference
    r = core.volume_profile_rolling_csr(np.array([10**9], np.int64), np.array([100.0]), np.array([99.9]), np.array([0, 2], np.int64),
                                        np.array([999, 1000], np.int32), np.array([1.0, 2.0], np.float32),
                                        np.array([0.5, 0.5], np.float32), 60.0, 27, 0.1, ctx=ctx)
    assert not r[0].any() and not r[3].any()
    # windows that reach back to the first bar, zero-volume levels, ties between the up and the down side of the POC
    nb = 40
    ts = np.arange(nb, dtype=np.int64) * 10**9
    rng = np.random.default_rng(4)
    lo = 1000 + rng.integers(-3, 3, nb)
    hi = lo + rng.integers(0, 6, nb)
    off = np.zeros(nb + 1, np.int64)
    off[1:] = np.cumsum(hi - lo + 1)
    lv = np.concatenate([np.arange(a, b + 1) for a, b in zip(lo, hi)]).astype(np.int32)
    bv = rng.choice([0.0, 1.0, 1.0, 2.0], len(lv)).astype(np.float32)
    sv = rng.choice([0.0, 1.0, 2.0], len(lv)).astype(np.float32)
    for nbins in (None, 3, 27):
        g = core.volume_profile_rolling_csr(ts, hi * 0.1, lo * 0.1, off, lv, bv, sv, 7.0, nbins, 0.1, ctx=ctx)
        o = oracle.volume_profile_rolling_csr(ts, hi * 0.1, lo * 0.1, off, lv, bv, sv, 7.0, nbins, 0.1)
        for q in range(4):
            assert_exact(g[q], o[q], f"nbins={nbins} out {q}")
