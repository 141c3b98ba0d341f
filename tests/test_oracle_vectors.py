"""CPU: the oracle against the reference's own hand-computed test vectors (tests/ref_vectors.py)."""
import types

import numpy as np
import pytest

import _vector_checks as chk
import oracle
import ref_vectors as V

API = types.SimpleNamespace(comp_bar_ohlcv=oracle.comp_bar_ohlcv, time_bar_indexer=oracle.time_bar_indexer,
                            comp_bar_directional_features=oracle.comp_bar_directional_features,
                            comp_bar_footprints=oracle.comp_bar_footprints, triple_barrier=oracle.triple_barrier,
                            average_uniqueness=oracle.average_uniqueness, return_attribution=oracle.return_attribution)


def test_ohlcv():
    chk.check_ohlcv(API)


def test_time_clock():
    chk.check_time_clock(API)


def test_directional():
    chk.check_directional(API)


def test_footprint():
    chk.check_footprint(API)


def test_tbm():
    chk.check_tbm(API)


@pytest.mark.parametrize("prices,want", V.TICK_SIZE)
def test_tick_size_host(prices, want):
    from finmlkit_b200.bar.utils import comp_price_tick_size
    assert comp_price_tick_size(np.array(prices)) == pytest.approx(want, rel=1e-9, abs=1e-15)


def test_tick_size_empty_raises():
    from finmlkit_b200.bar.utils import comp_price_tick_size
    with pytest.raises(ValueError):
        comp_price_tick_size(np.array([]))


def test_weights():
    chk.check_weights(API)


@pytest.mark.parametrize("u,last,want", V.TIME_DECAY)
def test_time_decay_host(u, last, want):
    from finmlkit_b200.label.weights import time_decay
    np.testing.assert_allclose(time_decay(np.array(u), last), np.array(want), rtol=1e-12)


def test_volume_profile_vectors():
    chk.check_volume_profile(oracle.volume_profile_rolling_csr)


@pytest.mark.parametrize("c", V.CLASS_BALANCE)
def test_class_balance_host(c):
    from finmlkit_b200.label.weights import class_balance_weights
    u, cw, sums, fw = class_balance_weights(np.array(c["labels"], np.int8), np.array(c["w"]))
    np.testing.assert_array_equal(u, np.array(c["uniq"], np.int8))
    np.testing.assert_allclose(sums, c["sums"], rtol=1e-12)
    np.testing.assert_allclose(cw, c["cw"], rtol=1e-12)
    idx = np.searchsorted(u, np.array(c["labels"]))
    np.testing.assert_allclose(fw, np.array(c["w"]) * np.array(c["cw"])[idx], rtol=1e-12)
