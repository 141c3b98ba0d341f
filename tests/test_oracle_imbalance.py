"""CPU: the imbalance / run bar oracle (OWN semantics, parity unpinned -- the reference only has stubs, logic.py:224-261)
against a pure-Python statement of the same definitions, and the host mirror's argument checks."""
import numpy as np
import pytest

import oracle


def _py_bars(b, thr, kind):
    idx, theta, nb, ns = [0], 0, 0, 0
    for i in range(1, len(b)):
        if kind == 0:
            theta += int(np.sign(b[i]))
            hit = abs(theta) >= thr
        else:
            nb += b[i] > 0
            ns += b[i] < 0
            hit = max(nb, ns) >= thr
        if hit:
            idx.append(i)
            theta = nb = ns = 0
    return np.array(idx, np.int64)


def _py_ema(b, et0, eb0, span, lo, hi):
    alpha = 2.0 / (span + 1.0)
    idx, thrs = [0], [np.nan]
    uT = vT = uB = vB = 0.0
    thr = min(max(et0 * abs(eb0), lo), hi)
    theta, last = 0, 0
    for i in range(1, len(b)):
        theta += int(np.sign(b[i]))
        if abs(theta) >= thr:
            idx.append(i); thrs.append(thr)
            T = float(i - last); mb = theta / T
            if vT == 0.0:
                uT, vT, uB, vB = T, 1.0, mb, 1.0
            else:
                uT = T + (1 - alpha) * uT; vT = 1.0 + (1 - alpha) * vT
                uB = mb + (1 - alpha) * uB; vB = 1.0 + (1 - alpha) * vB
            thr = min(max((uT / vT) * abs(uB / vB), lo), hi)
            theta, last = 0, i
    return np.array(idx, np.int64), np.array(thrs)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fixed_threshold_bars(seed):
    rng = np.random.default_rng(seed)
    b = rng.choice([-1, 0, 1], 5000, p=[0.45, 0.05, 0.5]).astype(np.int8)
    for thr in (1, 2.5, 7, 40):
        for kind in (0, 1):
            assert np.array_equal(oracle.imbalance_bar_indexer(b, thr, kind), _py_bars(b, thr, kind)), (thr, kind)
    assert np.array_equal(oracle.imbalance_bar_indexer(b[:1], 3, 0), [0])
    assert len(oracle.imbalance_bar_indexer(np.ones(50, np.int8), 1e9, 0)) == 1


def test_ema_adaptive_bars():
    rng = np.random.default_rng(7)
    b = np.where(np.cumsum(rng.random(20000) < 0.3) % 2 == 0, 1, -1).astype(np.int8)
    got = oracle.imbalance_bar_indexer_ema(b, 500.0, 0.05, 10, 5.0, 400.0)
    exp = _py_ema(b, 500.0, 0.05, 10, 5.0, 400.0)
    assert np.array_equal(got[0], exp[0])
    np.testing.assert_allclose(got[1][1:], exp[1][1:], rtol=1e-15)
    assert len(got[0]) > 10
