"""
Tick-level transforms of the sigma pipeline with the reference's call shape (feature/transforms.py:89-117 ``ReturnT``,
:308-332 ``EWMST``; ``Compose`` feature/kit.py:464-721), computed on the GPU.

* When the reference is importable (``finmlkit`` on ``sys.path``) ``ReturnT`` / ``EWMST`` SUBCLASS the reference's own
  classes -- they are ``CoreTransform``s (feature/base.py:141-251) and can sit inside a reference ``Feature`` / ``FeatureKit`` /
  ``Compose`` -- and only ``_nb`` / ``_pd`` are overridden; the Numba kernels of the reference are never called.  Otherwise
  small stand-ins with the same constructor, output naming and backend check are used.
* Device residency: called on a trades frame (``trades.data``) the transforms work on the frame's ONE device copy
  (``core.device_trades_for``); a ``Compose(ReturnT, EWMST)`` keeps the returns on the device and downloads only sigma; the
  sigma array is remembered together with its device copy (``core.register_device_copy``), so ``CUSUMBarKit(trades, sigma)``
  and ``TBMLabel`` do not upload it again.
"""
import numpy as np
import pandas as pd

from .. import core
from .core.utils import comp_lagged_returns
from .core.volatility import ewmst

try:                                   # the reference, when installed (never its kernels: only the transform base classes)
    from finmlkit.feature import transforms as _ref_t
    from finmlkit.feature import kit as _ref_kit
    if not hasattr(_ref_t, "SISOTransform"):
        _ref_t = None
except Exception:                      # noqa: BLE001 - any import problem means "not available"
    _ref_t = None
    _ref_kit = None


def _ts_ns(index: pd.Index) -> np.ndarray:
    # pandas 3 date_range is datetime64[us]: normalise to ns before taking int64
    if isinstance(index, pd.DatetimeIndex):
        return index.as_unit("ns").asi8
    return np.asarray(index.values).astype(np.int64)


def _trades_frame_device(x, input_col):
    """the shared device copy of ``x`` when ``x`` is a trades frame and ``input_col`` its price column, else None"""
    if not isinstance(x, pd.DataFrame) or input_col != "price" or not {"timestamp", "price", "amount"} <= set(x.columns):
        return None
    if len(x) == 0 or not isinstance(x.index, pd.DatetimeIndex):
        return None
    # the reference takes the timestamps from the index (feature/base.py:281-299); for a trades frame that is the timestamp
    # column (bar/data_model.py:186-190) -- check the ends before trusting the device column
    idx = x.index.as_unit("ns")
    tcol = x["timestamp"].values
    if int(idx[0].value) != int(tcol[0]) or int(idx[-1].value) != int(tcol[-1]):
        return None
    return core.device_trades_for(x, need_ts=True)


def _series_out(values, index, name):
    return pd.Series(values, index=index, name=name)


def _returns(x, input_col, window_sec, is_log, name, keep_on_device=False):
    tr = _trades_frame_device(x, input_col)
    if tr is not None:
        buf = core.lagged_returns_dev(tr, window_sec, is_log)
        if keep_on_device:
            return buf, tr
        out = _series_out(buf.download(np.float64, tr.n), x.index, name)
        core.register_device_copy(out, (buf, tr))
        return out
    s = x[input_col] if isinstance(x, pd.DataFrame) else x
    r = comp_lagged_returns(_ts_ns(s.index), s.values.astype(np.float64), window_sec, is_log)
    return _series_out(r, s.index, name)


def _ewm_std(x, input_col, half_life_sec, name):
    s = x[input_col] if isinstance(x, pd.DataFrame) else x
    hit = core.device_copy_of(s.values)
    if isinstance(hit, tuple):                      # (device buffer, trades handle) left by a device-resident ReturnT
        buf, tr = hit
        dev = core.ewmst_dev(tr, buf, half_life_sec)
        out = _series_out(dev.download(np.float64, tr.n), s.index, name)
        core.register_device_copy(out, dev)
        return out
    r = ewmst(_ts_ns(s.index), s.values.astype(np.float64), half_life_sec)
    return _series_out(r, s.index, name)


if _ref_t is not None:
    class ReturnT(_ref_t.ReturnT):
        """The reference's ``ReturnT`` (feature/transforms.py:89-117) with the GPU kernel behind ``_nb``."""

        def _nb(self, x):
            return _returns(x, self.requires[0], self.window_sec, self.is_log, self.output_name)

        def _pd(self, x):
            return self._nb(x)

    class EWMST(_ref_t.EWMST):
        """The reference's ``EWMST`` (feature/transforms.py:308-332) with the GPU kernel behind ``_nb``."""

        def _nb(self, x):
            return _ewm_std(x, self.requires[0], self.half_life_sec, self.output_name)

        def _pd(self, x):
            return self._nb(x)

    class Compose(_ref_kit.Compose):
        """The reference's ``Compose`` (feature/kit.py:464-721); a leading ``ReturnT`` -> ``EWMST`` chain on a trades frame
        runs device resident (only the final series is downloaded)."""

        def _run_pipeline(self, x, *, backend):
            fused = _fused_sigma(self.transforms, x, self.output_name)
            if fused is not None:
                return fused
            return super()._run_pipeline(x, backend=backend)
else:
    class _Transform:
        def __call__(self, x, *, backend="nb"):
            if backend not in ("nb", "pd"):
                raise ValueError(f"Unknown backend: {backend}")
            return self._nb(x)

    class ReturnT(_Transform):
        def __init__(self, window: pd.Timedelta = pd.Timedelta(seconds=1e-6), is_log: bool = False, input_col: str = "close"):
            self.window_sec = window.total_seconds() if isinstance(window, pd.Timedelta) else float(window)
            self.is_log, self.input_col = is_log, input_col
            self.output_name = f"{input_col}_ret{self.window_sec}s" if self.window_sec > 1e-6 else f"{input_col}_ret1"
            self.requires, self.produces = [input_col], [self.output_name]

        def _nb(self, x):
            return _returns(x, self.input_col, self.window_sec, self.is_log, self.output_name)

    class EWMST(_Transform):
        def __init__(self, half_life: pd.Timedelta, input_col: str = "y"):
            self.half_life_sec = half_life.total_seconds() if isinstance(half_life, pd.Timedelta) else float(half_life)
            self.input_col = input_col
            self.output_name = f"{input_col}_ewms{self.half_life_sec}s"
            self.requires, self.produces = [input_col], [self.output_name]

        def _nb(self, x):
            return _ewm_std(x, self.input_col, self.half_life_sec, self.output_name)

    class Compose(_Transform):
        """Chain of single-input transforms; output name is the chained name, e.g. ``price_ret3600.0s_ewms3600.0s``."""

        def __init__(self, *transforms):
            self.transforms = transforms
            first = transforms[0].output_name
            self.output_name = "_".join([first] + [t.output_name.split("_", 1)[1] if "_" in t.output_name else t.output_name
                                                   for t in transforms[1:]])

        def _nb(self, x):
            fused = _fused_sigma(self.transforms, x, self.output_name)
            if fused is not None:
                return fused
            out = self.transforms[0](x)
            for t in self.transforms[1:]:
                name = out.name
                t.input_col = name
                t.requires = [name]
                t.output_name = f"{name}_ewms{t.half_life_sec}s" if isinstance(t, EWMST) else f"{name}_ret{t.window_sec}s"
                out = t(out.to_frame())
            out.name = self.output_name
            return out


def _fused_sigma(transforms, x, final_name):
    """``Compose(ReturnT(..., input_col='price'), EWMST(...))`` on a trades frame: lagged returns and the EW std run back to
    back on the frame's device copy (fmk_lagged_returns_dev -> fmk_ewmst_dev) and ONLY sigma crosses PCIe.  None when the
    chain / input does not have that shape (the caller then runs the generic pipeline)."""
    if len(transforms) != 2 or not isinstance(transforms[0], ReturnT) or not isinstance(transforms[1], EWMST):
        return None
    t0, t1 = transforms
    if isinstance(x, pd.DataFrame) and final_name in x.columns:
        return None
    got = _returns(x, t0.requires[0], t0.window_sec, t0.is_log, None, keep_on_device=True) if _trades_frame_device(x, t0.requires[0]) is not None else None
    if got is None:
        return None
    buf, tr = got
    dev = core.ewmst_dev(tr, buf, t1.half_life_sec)
    del buf
    out = _series_out(dev.download(np.float64, tr.n), x.index, final_name)
    core.register_device_copy(out, dev)
    return out
