"""Order statistics inside the fused OHLCV kernel (csrc/reduce.cu k_bar_ohlcv_median): the staged shared-memory select,
its duplicate-heavy / all-equal shortcuts, the '(k+1)-th above the bucket' search and the generic fallback for bars longer
than the staging capacity -- against the oracle's np.median restatement (bit-exact: the median is a copy or the mean of
two copies of input values)."""
import numpy as np
import pytest

from helpers import assert_exact

pytestmark = pytest.mark.gpu


def _bars(rng, n, mean_len, long_every=0):
    cuts = [0]
    while cuts[-1] < n - 1:
        step = int(rng.integers(1, 2 * mean_len))
        if long_every and len(cuts) % long_every == 0:
            step = int(rng.integers(2049, 6000))
        cuts.append(min(cuts[-1] + step, n - 1))
    return np.array(cuts, np.int64)


def _amounts(rng, n, kind):
    if kind == "quantised":
        return np.round(rng.lognormal(-4, 1.2, n) + 0.001, 3)
    if kind == "continuous":
        return rng.lognormal(-4, 1.2, n)
    if kind == "few_values":
        return rng.choice([0.001, 0.002, 0.005, 0.01, 0.1, 0.25], n)
    if kind == "constant":
        return np.full(n, 0.125)
    if kind == "signed":      # negative and zero sizes: key order across the sign bit, +-0.0
        a = np.round(rng.normal(0, 1, n), 2)
        a[::17] = 0.0
        a[5::31] = -0.0
        return a
    if kind == "low_bits":    # values that differ only in the low 32 mantissa bits
        base = 0.1
        return base + rng.integers(0, 1000, n) * np.spacing(base)
    if kind == "wide":        # 300 orders of magnitude
        return 10.0 ** rng.uniform(-150, 150, n)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["quantised", "continuous", "few_values", "constant", "signed", "low_bits", "wide"])
@pytest.mark.parametrize("mean_len,long_every", [(12, 0), (700, 0), (300, 7)])
def test_median_matches_oracle(kind, mean_len, long_every, ctx):
    import oracle
    from finmlkit_b200 import core
    rng = np.random.default_rng(hash((kind, mean_len)) % 2**32)
    n = 400_000 if mean_len > 100 else 60_000
    px = np.round(30000 + np.cumsum(rng.normal(0, 1, n)), 1)
    qty = _amounts(rng, n, kind)
    idx = _bars(rng, n, mean_len, long_every)
    tr = core.DeviceTrades.upload(None, px, qty, ctx=ctx)
    ix = core.DeviceIndex.from_host(tr, idx)
    got = core.bar_ohlcv(tr, ix)
    exp = oracle.comp_bar_ohlcv(px, qty, idx)
    assert_exact(got[7], exp[7], f"median[{kind},{mean_len}]")
    for k in (0, 1, 2, 3, 6):
        assert_exact(got[k], exp[k], f"ohlcv[{k}]")
    if kind not in ("wide", "signed"):      # vwap = sum(pv)/sum(v): ill-conditioned for signed / 300-decade sizes
        np.testing.assert_allclose(got[5], exp[5], rtol=1e-9, atol=1e-300)


def test_median_tiny_counts(ctx):
    """Bars of 1..5 ticks (even counts need the (k+1)-th statistic from another bucket)."""
    import oracle
    from finmlkit_b200 import core
    rng = np.random.default_rng(2)
    n = 100_000
    px = np.full(n, 100.0)
    qty = np.round(rng.lognormal(-4, 1.2, n) + 0.001, 3)
    idx = np.concatenate([[0], np.cumsum(rng.integers(1, 6, 4000))]).astype(np.int64)
    idx = idx[idx < n]
    pad = np.arange(idx[-1] + 40, n, 40)           # keep n / n_bars >= 8 so the fused kernel is the one launched
    idx = np.concatenate([idx, pad]).astype(np.int64)
    tr = core.DeviceTrades.upload(None, px, qty, ctx=ctx)
    got = core.bar_ohlcv(tr, core.DeviceIndex.from_host(tr, idx))
    exp = oracle.comp_bar_ohlcv(px, qty, idx)
    assert_exact(got[7], exp[7], "median tiny")
