"""
Host-side synthetic trade stream (SURVEY.md section 8d shape): BTCUSDT-like ticks.

``synth_trades`` is the small-size generator used by tests and golden fixtures.  The 1e9-tick bench streams are produced
on the device by ``fmk_synth_trades`` (csrc/synth.cu) with the same distributional shape (it is NOT bit-identical to this
numpy generator; parity at bench size is checked by copying the device stream back to the host oracle).
"""
import numpy as np


def synth_trades(n: int, seed: int = 42, p0: float = 30000.0):
    """Returns (ts int64 ns, price f64 rounded to 0.1, amount f64 rounded to 0.001, side int8 +-1)."""
    rng = np.random.default_rng(seed)
    gaps = 1.0 + np.floor(rng.exponential(50e6, n))                 # ns, mean 50 ms
    ts = 1_700_000_000_000_000_000 + np.cumsum(gaps).astype(np.int64)
    ts = ts // 1_000_000 * 1_000_000                                # floor to ms -> duplicate timestamps
    px = np.round(p0 * np.exp(np.cumsum(rng.normal(0.0, 2e-5, n))), 1)
    qty = np.round(rng.lognormal(-4.0, 1.2, n) + 0.001, 3)
    flips = rng.random(n) < 0.3
    side = np.where(np.cumsum(flips) % 2 == 0, 1, -1).astype(np.int8)
    return ts, px, qty, side
