"""Rolling volume profile (feature/core/volume.py:133-456 of the reference): POC / HVA / LVA are integer
price levels decided on float32 sums accumulated in the reference's order -> bit-exact; the fraction above the POC is a
float32 of a float64 quotient -> bit-exact as well.  Fixtures: tests/golden/volprofile.npz (generated from the imported
reference by tests/golden/make_golden.py --only-volprofile)."""
import numpy as np
import pytest

from helpers import assert_exact, load_case


def _cases(g):
    return [(float(w), int(nb) if nb else None) for w, nb in g["cases"]]


def test_oracle_volprofile_golden():
    import oracle
    g = load_case("volprofile")
    for k, (w, nbins) in enumerate(_cases(g)):
        r = oracle.volume_profile_rolling_csr(g["ts"], g["high"], g["low"], g["off"], g["levels"], g["buy"], g["sell"], w, nbins, 0.1)
        for q in range(4):
            assert_exact(r[q], g[f"ref_{k}_{q}"], f"case {k} ({w}, {nbins}) out {q}")
    n = 64
    off = np.arange(n + 1, dtype=np.int64)
    r = oracle.volume_profile_rolling_csr(np.arange(n, dtype=np.int64) * 10**9, np.full(n, 100.0), np.full(n, 100.0), off,
                                          np.full(n, 1000, np.int32), (1.0 + np.arange(n)).astype(np.float32),
                                          np.full(n, 0.5, np.float32), 5.0, 27, 0.1)
    for q in range(4):
        assert_exact(r[q], g[f"flat_ref_{q}"], f"flat out {q}")


@pytest.mark.gpu
def test_gpu_volprofile_golden(ctx):
    from finmlkit_b200 import core
    g = load_case("volprofile")
    for k, (w, nbins) in enumerate(_cases(g)):
        r = core.volume_profile_rolling_csr(g["ts"], g["high"], g["low"], g["off"], g["levels"], g["buy"], g["sell"], w, nbins, 0.1, ctx=ctx)
        assert r[0].dtype == np.int32 and r[3].dtype == np.float32
        for q in range(4):
            assert_exact(r[q], g[f"ref_{k}_{q}"], f"case {k} ({w}, {nbins}) out {q}")
    n = 64
    off = np.arange(n + 1, dtype=np.int64)
    r = core.volume_profile_rolling_csr(np.arange(n, dtype=np.int64) * 10**9, np.full(n, 100.0), np.full(n, 100.0), off,
                                        np.full(n, 1000, np.int32), (1.0 + np.arange(n)).astype(np.float32),
                                        np.full(n, 0.5, np.float32), 5.0, 27, 0.1, ctx=ctx)
    for q in range(4):
        assert_exact(r[q], g[f"flat_ref_{q}"], f"flat out {q}")


@pytest.mark.gpu
def test_gpu_volumepro_on_kit_footprints(ctx):
    """DollarBarKit.build_footprints -> VolumePro.compute (the reference's call chain, core/volume.py:49-86), against the
    oracle on a 300k-tick stream, plus a non-contiguous level list (binary-search path)."""
    import pandas as pd
    import oracle
    from finmlkit_b200.bar.data_model import TradesData
    from finmlkit_b200.bar.kit import DollarBarKit
    from finmlkit_b200.feature.core.volume import VolumePro, volume_profile_rolling
    from finmlkit_b200.synth import synth_trades
    ts, px, qty, side = synth_trades(300_000, seed=12)
    kit = DollarBarKit(TradesData(ts, px, qty, side=side), 5e4)
    bars = kit.build_ohlcv()
    fp = kit.build_footprints(price_tick_size=0.1)
    vp = VolumePro(pd.Timedelta(seconds=120), n_bins=27)
    poc, hva, lva, pct = vp.compute(bars, fp)
    off = np.zeros(len(fp) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in fp.price_levels])
    cat = lambda xs, dt: np.concatenate([np.asarray(x, dt) for x in xs])   # noqa: E731
    o = oracle.volume_profile_rolling_csr(fp.bar_timestamps, bars.high.values, bars.low.values, off, cat(fp.price_levels, np.int32),
                                          cat(fp.buy_volumes, np.float32), cat(fp.sell_volumes, np.float32), 120.0, 27, 0.1)
    for got, exp in ((poc, o[0]), (hva, o[1]), (lva, o[2])):
        e = exp * 0.1
        assert_exact(got, np.where(e == 0, np.nan, e), "VolumePro price")
    assert_exact(pct, o[3], "pct above poc")
    assert np.isnan(poc[0]) and np.isfinite(poc[-1])
    # drop the middle level of every wide bar: levels no longer contiguous
    pl, bl, sl = [], [], []
    for a, b, c in zip(fp.price_levels, fp.buy_volumes, fp.sell_volumes):
        a, b, c = np.asarray(a), np.asarray(b), np.asarray(c)
        if len(a) >= 3:
            keep = np.ones(len(a), bool); keep[len(a) // 2] = False
            a, b, c = a[keep], b[keep], c[keep]
        pl.append(a); bl.append(b); sl.append(c)
    r = volume_profile_rolling(fp.bar_timestamps, bars.high.values, bars.low.values, pl, bl, sl, 60.0, None, 0.1, ctx=ctx)
    off2 = np.zeros(len(pl) + 1, np.int64)
    off2[1:] = np.cumsum([len(x) for x in pl])
    o2 = oracle.volume_profile_rolling_csr(fp.bar_timestamps, bars.high.values, bars.low.values, off2, cat(pl, np.int32),
                                           cat(bl, np.float32), cat(sl, np.float32), 60.0, None, 0.1)
    for q in range(4):
        assert_exact(r[q], o2[q], f"non-contiguous out {q}")
