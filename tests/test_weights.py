"""Sample weights on ticks (label/weights.py of the reference).

CPU: the oracle against the fixtures generated from the imported reference (tests/golden/weights.npz) and the host-side
time-decay / class-balance arithmetic.  GPU: the CUDA path (csrc/weights.cu) through the C ABI against the fixtures and,
on a larger stream, against the oracle.  Tolerances: concurrency bit-exact (int16, wrap-around included); float64 weights
1e-9 relative (north_star), return attribution additionally abs 1e-12 x sum|terms| (a signed sum: reordering error scales
with the terms, not with a cancelled result)."""
import numpy as np
import pytest

from helpers import assert_exact, assert_f64, load_case


def _g():
    return load_case("weights")


def test_oracle_weights_golden():
    import oracle
    g = _g()
    w, c = oracle.average_uniqueness(g["a_ts"], g["a_ev"], g["a_touch"])
    assert_exact(c, g["a_ref_conc"], "conc")
    assert_exact(w, g["a_ref_avg_u"], "avg_u")
    assert_exact(oracle.return_attribution(g["a_ev"], g["a_touch"], g["a_px"], c, False), g["a_ref_ra"], "ra")
    assert_f64(oracle.return_attribution(g["a_ev"], g["a_touch"], g["a_px"], c, True), g["a_ref_ra_norm"], "ra norm", rtol=1e-12)
    w, c = oracle.average_uniqueness(g["b_ts"], g["b_ev"], g["b_touch"])
    assert_exact(c, g["b_ref_conc"], "wrap conc")
    assert_exact(w, g["b_ref_avg_u"], "wrap avg_u")
    assert_exact(oracle.return_attribution(g["b_ev"], g["b_touch"], g["b_px"], c, False), g["b_ref_ra"], "wrap ra")
    w, c = oracle.average_uniqueness(g["b_ts"], g["c_ev"], g["c_touch"])
    assert_exact(c, g["c_ref_conc"], "zero-wrap conc")
    assert_exact(w, g["c_ref_avg_u"], "zero-wrap avg_u")
    assert_exact(oracle.return_attribution(g["c_ev"], g["c_touch"], g["b_px"], c, False), g["c_ref_ra"], "zero-wrap ra")


def test_host_decay_and_class_balance_golden():
    from finmlkit_b200.label.weights import class_balance_weights, time_decay
    g = _g()
    assert_exact(time_decay(g["a_ref_avg_u"], 0.5), g["a_ref_decay_05"], "decay 0.5")
    assert_exact(time_decay(g["a_ref_avg_u"], -0.3), g["a_ref_decay_m03"], "decay -0.3")
    cb = class_balance_weights(g["a_labels"], g["a_ref_avg_u"])
    assert_exact(cb[0], g["a_ref_cb_0"], "classes")
    for k in (1, 2, 3):
        assert_f64(cb[k], g[f"a_ref_cb_{k}"], f"cb[{k}]", rtol=1e-13)
    with pytest.raises(ValueError, match="last_weight must lie"):
        time_decay(g["a_ref_avg_u"], 1.5)
    with pytest.raises(ValueError, match="grater than 0"):
        time_decay(np.zeros(4), 0.5)


def test_oracle_length_mismatch():
    import oracle
    with pytest.raises(ValueError, match="same length"):
        oracle.average_uniqueness(np.arange(10), np.array([1, 2]), np.array([3]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_gpu_weights_golden(case, ctx):
    from finmlkit_b200.label import weights as W
    g = _g()
    ts = g["a_ts"] if case == "a" else g["b_ts"]
    px = g["a_px"] if case == "a" else g["b_px"]
    ev, tc = g[f"{case}_ev"], g[f"{case}_touch"]
    w, c = W.average_uniqueness(ts, ev, tc, ctx=ctx)
    assert c.dtype == np.int16
    assert_exact(c, g[f"{case}_ref_conc"], "conc")
    assert_f64(w, g[f"{case}_ref_avg_u"], "avg_u")
    ra = W.return_attribution(ev, tc, px, c, False, ctx=ctx)
    assert_f64(ra, g[f"{case}_ref_ra"], "ra", atol=1e-12)
    if case == "a":
        assert_f64(W.return_attribution(ev, tc, px, c, True, ctx=ctx), g["a_ref_ra_norm"], "ra norm", atol=1e-12)


@pytest.mark.gpu
def test_gpu_weights_vs_oracle_large(ctx):
    """2M ticks, 20k overlapping labels with long paths (tile-prefix path), device-resident price column."""
    import oracle
    from finmlkit_b200 import core
    from finmlkit_b200.synth import synth_trades
    n = 2_000_000
    ts, px, qty, side = synth_trades(n, seed=9)
    rng = np.random.default_rng(1)
    ev = np.sort(rng.integers(0, n - 1, 20000)).astype(np.int64)
    tc = np.minimum(ev + rng.integers(0, 60000, len(ev)), n - 1).astype(np.int64)
    tc[::97] = ev[::97]                       # single-tick labels
    tc[5] = ev[5] - 1 if ev[5] > 0 else 0     # empty label
    ow, oc = oracle.average_uniqueness(ts, ev, tc)
    ora = oracle.return_attribution(ev, tc, px, oc, False)
    tr = core.DeviceTrades.upload(None, px, qty, ctx=ctx)
    u, r, c = core.sample_weights_dev(tr, ev, tc, want_concurrency=True)
    assert_exact(c, oc, "conc")
    assert_f64(u, ow, "avg_u")
    assert_f64(r, ora, "ra", atol=1e-11)
    un, rn = core.sample_weights_dev(tr, ev, tc, normalize=True)
    assert_f64(rn, oracle.return_attribution(ev, tc, px, oc, True), "ra norm", atol=1e-9)
    assert abs(rn.sum() - len(ev)) < 1e-6


@pytest.mark.gpu
def test_gpu_weights_errors(ctx):
    from finmlkit_b200.label import weights as W
    with pytest.raises(ValueError, match="same length"):
        W.average_uniqueness(np.arange(10), np.array([1, 2]), np.array([3]), ctx=ctx)
    with pytest.raises(ValueError, match=r"must lie in \[0"):
        W.average_uniqueness(np.arange(10), np.array([1, 2]), np.array([3, 10]), ctx=ctx)
    with pytest.raises(ValueError, match="cannot normalize"):
        W.return_attribution(np.array([1]), np.array([3]), np.ones(10), np.ones(10, np.int16), True, ctx=ctx)


@pytest.mark.gpu
def test_gpu_sample_weights_kit(ctx):
    """TBMLabel.compute_weights / SampleWeights.compute_info_weights against the oracle on the golden TBM events."""
    import pandas as pd
    import oracle
    from finmlkit_b200.bar.data_model import TradesData
    from finmlkit_b200.label.kit import SampleWeights
    g = _g()
    td = TradesData(g["a_ts"], g["a_px"], np.ones(len(g["a_ts"])), side=np.ones(len(g["a_ts"]), np.int8))
    lab = pd.DataFrame({"event_idx": g["a_ev"], "touch_idx": g["a_touch"]}, index=pd.to_datetime(g["a_ts"][g["a_ev"]]))
    out = SampleWeights.compute_info_weights(td, lab)
    assert list(out.columns) == ["avg_uniqueness", "return_attribution"]
    assert_f64(out["avg_uniqueness"].values, g["a_ref_avg_u"], "kit avg_u")
    assert_f64(out["return_attribution"].values, g["a_ref_ra"], "kit ra", atol=1e-12)
    fin = SampleWeights.compute_final_weights(out["avg_uniqueness"], 0.5, out["return_attribution"], labels=pd.Series(g["a_labels"], index=out.index))
    assert list(fin.columns) == ["time_decay_weights", "return_attribution", "weights"]
    assert abs(fin["return_attribution"].sum() - len(out)) < 1e-9
