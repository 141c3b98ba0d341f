"""GPU: the CUDA path, called through the mirrored reference API, against the reference's hand-computed vectors."""
import types

import pytest

import _vector_checks as chk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from finmlkit_b200.bar import base, logic
    from finmlkit_b200.label import tbm, weights
    return types.SimpleNamespace(comp_bar_ohlcv=base.comp_bar_ohlcv, time_bar_indexer=logic._time_bar_indexer,
                                 comp_bar_directional_features=base.comp_bar_directional_features,
                                 comp_bar_footprints=base.comp_bar_footprints, triple_barrier=tbm.triple_barrier,
                                 average_uniqueness=weights.average_uniqueness, return_attribution=weights.return_attribution)


def test_ohlcv(api):
    chk.check_ohlcv(api)


def test_time_clock(api):
    chk.check_time_clock(api)


def test_directional(api):
    chk.check_directional(api)


def test_footprint(api):
    chk.check_footprint(api)


def test_tbm(api):
    chk.check_tbm(api)


def test_weights(api):
    chk.check_weights(api)


def test_volume_profile_vectors(ctx):
    from finmlkit_b200 import core
    chk.check_volume_profile(lambda *a: core.volume_profile_rolling_csr(*a, ctx=ctx))
