"""
Drop-in installation into an importable reference: ``install()`` rebinds the reference's hot-path
functions and bar kits to the GPU implementations of this package, everywhere the reference has already bound them.

The reference (quantscious/finmlkit) is pure Python + Numba and reaches its hot path through module globals
(``finmlkit.bar.base.comp_bar_ohlcv`` called by ``BarBuilderBase.build_ohlcv`` -- bar/base.py:147 --,
``finmlkit.feature.transforms`` importing ``vpin`` / ``realized_vol`` / ``ewmst`` / ``comp_lagged_returns`` from
``feature/core``, ``TBMLabel`` calling ``triple_barrier`` -- label/kit.py:290 -- ...).  Rebinding those names is therefore
the whole integration: the feature framework (``Feature`` / ``Compose`` / ``FeatureKit``, feature/kit.py), ``TBMLabel``,
``SampleWeights`` and user code keep working unchanged and the tick-level loops run on the device.  ``uninstall()`` restores
the Numba functions.  tests/test_gpu_ref_suite.py runs the reference's OWN test files under ``install()``.
"""
import importlib
import sys

_SAVED = []          # (module, name, original object)


def _targets():
    from .bar import base as b_base, kit as b_kit, logic as b_logic, utils as b_utils
    from .feature.core import utils as f_utils, volatility as f_vol, volume as f_volume
    from .label import tbm as l_tbm, weights as l_w
    from .sampling import filters as s_f
    return {
        "finmlkit.bar.logic": {n: getattr(b_logic, n) for n in ("_time_bar_indexer", "_tick_bar_indexer", "_volume_bar_indexer",
                                                                "_dollar_bar_indexer", "_cusum_bar_indexer")},
        "finmlkit.bar.base": {n: getattr(b_base, n) for n in ("comp_bar_ohlcv", "comp_bar_directional_features",
                                                               "comp_bar_trade_size_features", "comp_bar_footprints",
                                                               "comp_footprint_features")},
        "finmlkit.bar.kit": {n: getattr(b_kit, n) for n in ("TimeBarKit", "TickBarKit", "VolumeBarKit", "DollarBarKit", "CUSUMBarKit")},
        "finmlkit.bar.utils": {n: getattr(b_utils, n) for n in ("comp_trade_side_vector", "merge_split_trades")},
        "finmlkit.feature.core.utils": {"comp_lagged_returns": f_utils.comp_lagged_returns},
        "finmlkit.feature.core.volatility": {n: getattr(f_vol, n) for n in ("ewmst", "ewms", "realized_vol")},
        "finmlkit.feature.core.volume": {n: getattr(f_volume, n) for n in ("vpin", "comp_flow_acceleration", "volume_profile_rolling")},
        "finmlkit.label.tbm": {"triple_barrier": l_tbm.triple_barrier},
        "finmlkit.label.weights": {n: getattr(l_w, n) for n in ("average_uniqueness", "return_attribution")},
        "finmlkit.sampling.filters": {"cusum_filter": s_f.cusum_filter},
    }


def install(verbose: bool = False) -> int:
    """Rebind the reference's hot-path names to the GPU implementations; returns the number of bindings replaced.
    Requires ``finmlkit`` to be importable (e.g. ``baseline/_ref`` or a site-packages install on ``sys.path``)."""
    if _SAVED:
        return 0
    import finmlkit  # noqa: F401  (the reference; ImportError here means there is nothing to drop into)
    # make sure every module that binds hot-path names at import time is loaded before the scan
    for m in ("finmlkit.bar.kit", "finmlkit.bar.base", "finmlkit.bar.logic", "finmlkit.bar.utils", "finmlkit.feature.transforms",
              "finmlkit.feature.core.volume", "finmlkit.feature.core.volatility", "finmlkit.feature.core.utils",
              "finmlkit.label.kit", "finmlkit.label.tbm", "finmlkit.label.weights", "finmlkit.sampling.filters"):
        try:
            importlib.import_module(m)
        except Exception as e:      # optional dependencies of the reference (PyTables ...) may be missing
            if verbose:
                print(f"dropin: {m} not importable: {e}", file=sys.stderr)
    replaced = 0
    for modname, names in _targets().items():
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for name, new in names.items():
            old = getattr(mod, name, None)
            if old is None or old is new:
                continue
            # every loaded reference module that holds the same object under any name (`from .core.volume import vpin`)
            for m2name, m2 in list(sys.modules.items()):
                if m2 is None or not (m2name == "finmlkit" or m2name.startswith("finmlkit.")):
                    continue
                for attr, val in list(vars(m2).items()):
                    if val is old:
                        _SAVED.append((m2, attr, old))
                        setattr(m2, attr, new)
                        replaced += 1
    if verbose:
        print(f"dropin: {replaced} bindings now run on the GPU", file=sys.stderr)
    return replaced


def uninstall() -> None:
    while _SAVED:
        mod, attr, old = _SAVED.pop()
        setattr(mod, attr, old)
