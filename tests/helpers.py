import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STREAM_CASES = ["synth_20k", "adversarial_6k", "giant_4k"]
BAR_KINDS = ["time", "dollar", "volume", "tick", "cusum"]

OHLCV_NAMES = ["open", "high", "low", "close", "volume", "vwap", "trades", "median"]


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def clock_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "clock_*.npz")))


def assert_exact(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        same = (a == b) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    if not same.all():
        k = int(np.argmin(same))
        raise AssertionError(f"{what}: {int((~same).sum())} mismatches, first at {k}: {a[k]!r} vs {b[k]!r}")


def assert_f64(a, b, what="", rtol=1e-9, atol=1e-12):
    """float64 outputs: <= 1e-9 relative (north_star tolerance), abs 1e-12 near zero; NaN/inf must coincide."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    fin = np.isfinite(a) & np.isfinite(b)
    nonfin_same = (np.isnan(a) & np.isnan(b)) | (a == b)
    assert (fin | nonfin_same).all(), f"{what}: non-finite pattern differs"
    err = np.abs(a[fin] - b[fin])
    tol = atol + rtol * np.abs(b[fin])
    if (err > tol).any():
        k = int(np.argmax(err - tol))
        raise AssertionError(f"{what}: max violation {err[k]:.3e} (tol {tol[k]:.3e}) values {a[fin][k]!r} vs {b[fin][k]!r}")


def assert_f32_ulp(a, b, what="", ulps=1, atol=0.0):
    """float32 outputs the reference computes in float64 and casts: within `ulps` float32 ulp."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    fin = np.isfinite(a) & np.isfinite(b)
    nonfin_same = (np.isnan(a) & np.isnan(b)) | (a == b)
    assert (fin | nonfin_same).all(), f"{what}: non-finite pattern differs"
    err = np.abs(a[fin].astype(np.float64) - b[fin].astype(np.float64))
    tol = ulps * np.spacing(np.maximum(np.abs(a[fin]), np.abs(b[fin]))).astype(np.float64) + atol
    if (err > tol).any():
        k = int(np.argmax(err - tol))
        raise AssertionError(f"{what}: {err[k]:.3e} > {tol[k]:.3e}: {a[fin][k]!r} vs {b[fin][k]!r}")


def check_ohlcv(got, ref, what):
    for k, name in enumerate(OHLCV_NAMES):
        w = f"{what}.ohlcv.{name}"
        if name in ("open", "high", "low", "close", "trades"):
            assert_exact(got[k], ref[k], w)
        elif name == "volume":
            assert_f32_ulp(got[k], ref[k], w)
        elif name == "median":
            assert_exact(got[k], ref[k], w)        # order statistics are exact
        else:
            assert_f64(got[k], ref[k], w)


DIR_INT = {0, 1, 8, 9}


def check_directional(got, ref, what):
    for k in range(14):
        w = f"{what}.dir[{k}]"
        if k in DIR_INT:
            assert_exact(got[k], ref[k], w)
        elif k >= 10:
            # running signed sums return to ~0 by cancellation: their float64 rounding noise (1e-16 x flow size) is
            # order-dependent, so the min/max need an absolute floor next to the 1-ulp float32 tolerance
            scale = max(float(np.max(np.abs(ref[k][np.abs(ref[k]) < 1e8]), initial=1.0)), 1.0)
            assert_f32_ulp(got[k], ref[k], w, atol=1e-12 * scale)
        else:
            assert_f32_ulp(got[k], ref[k], w)


def check_trade_size(got, ref, what):
    for k in range(4):
        assert_f32_ulp(got[k], ref[k], f"{what}.tsize[{k}]", ulps=2)


def check_footprint_csr(got, ref_off, ref, levels_scale, what):
    """got: CSR tuple from core.bar_footprints_csr / oracle; ref: list of 13 reference arrays (7 flat + 6 per bar)."""
    assert_exact(got[0], ref_off, f"{what}.fp.offsets")
    for k in range(7):
        assert_exact(got[1 + k], ref[k], f"{what}.fp[{k}]")   # levels, f32 volumes (order-preserving), ticks, flags
    for k in range(7, 11):
        assert_exact(got[1 + k], ref[k], f"{what}.fp[{k}]")
    # vp_skew is float32 rounding noise around 0: absolute tolerance scaled by the level magnitude
    assert_f64(got[12], ref[11], f"{what}.fp.vp_skew", rtol=0, atol=4e-6 * levels_scale + 1e-6)
    assert_f64(got[13], ref[12], f"{what}.fp.vp_gini", rtol=0, atol=2e-6)
