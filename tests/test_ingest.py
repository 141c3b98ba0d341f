"""Event sampler and ingest scans: cusum_filter (sampling/filters.py), comp_trade_side_vector and
merge_split_trades (bar/utils.py).  CPU: oracle vs fixtures generated from the imported reference (tests/golden/ingest.npz).
GPU: CUDA path through the mirrored API vs the fixtures and, at 2e6 ticks, vs the oracle.  Everything here is integer /
index / copy work except the merged float32 amounts (sequential float32 sums in arrival order): all bit-exact."""
import numpy as np
import pytest

from helpers import assert_exact, load_case


def _g():
    return load_case("ingest")


def test_oracle_ingest_golden():
    import oracle
    g = _g()
    assert_exact(oracle.cusum_filter(g["f_px"], g["f_thr_const"]), g["f_ref_const"], "filter const")
    assert_exact(oracle.cusum_filter(g["f_px"], g["f_thr_arr"]), g["f_ref_arr"], "filter arr")
    assert_exact(oracle.comp_trade_side_vector(g["s_px"]), g["s_ref"], "tick rule")
    r = oracle.merge_split_trades(g["m_ts"], g["m_px"], g["m_am"], g["m_ibm"])
    for k in range(4):
        assert_exact(r[k], g[f"m_ref_{k}"], f"merge[{k}]")
    r = oracle.merge_split_trades(g["m_ts"], g["m_px"], g["m_am"], None)
    for k in range(3):
        assert_exact(r[k], g[f"m_ref_noside_{k}"], f"merge noside[{k}]")
    assert len(r[3]) == 0


def test_oracle_filter_errors():
    import oracle
    with pytest.raises(ValueError, match="at least 2 elements"):
        oracle.cusum_filter(np.array([1.0]), np.array([0.1]))
    with pytest.raises(ValueError, match="Threshold array must either"):
        oracle.cusum_filter(np.array([1.0, 2.0, 3.0]), np.array([0.1, 0.2]))


@pytest.mark.gpu
def test_gpu_ingest_golden(ctx):
    from finmlkit_b200.bar.utils import comp_trade_side_vector, merge_split_trades
    from finmlkit_b200.sampling.filters import cusum_filter
    g = _g()
    ev = cusum_filter(g["f_px"], g["f_thr_const"], ctx=ctx)
    assert ev.dtype == np.int64
    assert_exact(ev, g["f_ref_const"], "filter const")
    assert_exact(cusum_filter(g["f_px"], g["f_thr_arr"], ctx=ctx), g["f_ref_arr"], "filter arr")
    s = comp_trade_side_vector(g["s_px"], ctx=ctx)
    assert s.dtype == np.int8
    assert_exact(s, g["s_ref"], "tick rule")
    r = merge_split_trades(g["m_ts"], g["m_px"], g["m_am"], g["m_ibm"], ctx=ctx)
    for k in range(4):
        assert_exact(r[k], g[f"m_ref_{k}"], f"merge[{k}]")
    assert r[2].dtype == np.float32 and r[3].dtype == np.int8
    r = merge_split_trades(g["m_ts"], g["m_px"], g["m_am"], None, ctx=ctx)
    for k in range(3):
        assert_exact(r[k], g[f"m_ref_noside_{k}"], f"merge noside[{k}]")
    assert len(r[3]) == 0
    with pytest.raises(ValueError, match="at least 2 elements"):
        cusum_filter(np.array([1.0]), np.array([0.1]), ctx=ctx)
    with pytest.raises(ValueError, match="Threshold array must either"):
        cusum_filter(np.array([1.0, 2.0, 3.0]), np.array([0.1, 0.2]), ctx=ctx)


@pytest.mark.gpu
def test_gpu_ingest_vs_oracle_2e6(ctx, monkeypatch):
    import oracle
    from finmlkit_b200.bar.utils import comp_trade_side_vector, merge_split_trades
    from finmlkit_b200.sampling.filters import cusum_filter
    from finmlkit_b200.synth import synth_trades
    n = 2_000_000
    ts, px, qty, side = synth_trades(n, seed=31)
    for thr in (np.array([5e-4]), np.array([2e-5])):
        assert_exact(cusum_filter(px, thr, ctx=ctx), oracle.cusum_filter(px, thr), f"filter {thr}")
    monkeypatch.setenv("FMK_CUSUM_CH", "64")           # thousands of chunks, many fix-point rounds
    assert_exact(cusum_filter(px[:200_000], np.array([5e-4]), ctx=ctx), oracle.cusum_filter(px[:200_000], np.array([5e-4])), "filter small chunks")
    monkeypatch.delenv("FMK_CUSUM_CH")
    assert_exact(comp_trade_side_vector(px, ctx=ctx), oracle.comp_trade_side_vector(px), "tick rule")
    # synthetic timestamps are floored to ms -> real duplicate-timestamp runs; sort by (ts, price, side) like the ingest does
    ibm = side < 0
    order = np.lexsort((ibm, px, ts))
    a32 = qty.astype(np.float32)
    got = merge_split_trades(ts[order], px[order], a32[order], ibm[order], ctx=ctx)
    exp = oracle.merge_split_trades(ts[order], px[order], a32[order], ibm[order])
    for k in range(4):
        assert_exact(got[k], exp[k], f"merge[{k}]")
    assert len(got[0]) < n
