"""
Tick-level transforms of the sigma pipeline with the reference's call shape (feature/transforms.py:89-117 ``ReturnT``,
:308-332 ``EWMST``; ``Compose`` naming feature/kit.py:637-641).  Only the ``"nb"`` role (here: GPU) exists; like
``CoreTransform.__call__`` (feature/base.py:247-251) any other backend string raises ``ValueError``.
"""
import numpy as np
import pandas as pd

from .core.utils import comp_lagged_returns
from .core.volatility import ewmst


def _ts_ns(index: pd.Index) -> np.ndarray:
    # pandas 3 date_range is datetime64[us]: normalise to ns before taking int64
    if isinstance(index, pd.DatetimeIndex):
        return index.as_unit("ns").asi8
    return np.asarray(index.values).astype(np.int64)


class _Transform:
    def __call__(self, x, *, backend="nb"):
        if backend not in ("nb", "pd"):
            raise ValueError(f"Unknown backend: {backend}")
        return self._nb(x)


class ReturnT(_Transform):
    def __init__(self, window: pd.Timedelta, is_log: bool = False, input_col: str = "close"):
        self.window_sec = window.total_seconds() if isinstance(window, pd.Timedelta) else float(window)
        self.is_log, self.input_col = is_log, input_col
        self.output_name = f"{input_col}_ret{self.window_sec}s"

    def _nb(self, x):
        s = x[self.input_col] if isinstance(x, pd.DataFrame) else x
        r = comp_lagged_returns(_ts_ns(s.index), s.values.astype(np.float64), self.window_sec, self.is_log)
        return pd.Series(r, index=s.index, name=self.output_name)


class EWMST(_Transform):
    def __init__(self, half_life: pd.Timedelta, input_col: str = "close"):
        self.half_life_sec = half_life.total_seconds() if isinstance(half_life, pd.Timedelta) else float(half_life)
        self.input_col = input_col
        self.output_name = f"{input_col}_ewms{self.half_life_sec}s"

    def _nb(self, x):
        s = x[self.input_col] if isinstance(x, pd.DataFrame) else x
        r = ewmst(_ts_ns(s.index), s.values.astype(np.float64), self.half_life_sec)
        return pd.Series(r, index=s.index, name=self.output_name)


class Compose(_Transform):
    """Chain of single-input transforms; output name is the chained name, e.g. ``price_ret3600.0s_ewms3600.0s``."""

    def __init__(self, *transforms):
        self.transforms = transforms

    def _nb(self, x):
        out = self.transforms[0](x)
        for t in self.transforms[1:]:
            name = out.name
            t.input_col = name
            t.output_name = f"{name}_ewms{t.half_life_sec}s" if isinstance(t, EWMST) else f"{name}_ret{t.window_sec}s"
            out = t(out.to_frame())
        return out
