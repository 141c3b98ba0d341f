"""Bar kits with the reference's constructors (finmlkit/bar/kit.py:12-181); indices are computed on the device."""
from typing import Tuple

import numpy as np
import pandas as pd

from .. import core
from .base import BarBuilderBase


class TimeBarKit(BarBuilderBase):
    """kit.py:12-35."""

    _needs_device_ts = True

    def __init__(self, trades, period: pd.Timedelta, ctx=None):
        super().__init__(trades, ctx)
        self.interval = period.total_seconds()

    def _comp_bar_close(self) -> Tuple[np.ndarray, np.ndarray]:
        self._dev_index = core.time_bar_index(self._device(), self.interval)
        return self._dev_index.download()


class TickBarKit(BarBuilderBase):
    """kit.py:38-69."""

    def __init__(self, trades, tick_count_thrs: int, ctx=None):
        super().__init__(trades, ctx)
        self.tick_count_thrs = tick_count_thrs

    def _comp_bar_close(self):
        self._dev_index = core.tick_bar_index(self._device(), self.tick_count_thrs)
        return self._download_index()


class VolumeBarKit(BarBuilderBase):
    """kit.py:72-104."""

    def __init__(self, trades, volume_ths: float, ctx=None):
        super().__init__(trades, ctx)
        self.volume_ths = volume_ths

    def _comp_bar_close(self):
        self._dev_index = core.volume_bar_index(self._device(), self.volume_ths)
        return self._download_index()


class DollarBarKit(BarBuilderBase):
    """kit.py:107-138."""

    def __init__(self, trades, dollar_thrs: float, ctx=None):
        super().__init__(trades, ctx)
        self.dollar_thrs = dollar_thrs

    def _comp_bar_close(self):
        self._dev_index = core.dollar_bar_index(self._device(), self.dollar_thrs)
        return self._download_index()


class CUSUMBarKit(BarBuilderBase):
    """kit.py:141-181.  Like the reference, NaNs of ``sigma`` are forward-filled in place by the indexer."""

    _needs_device_ts = True

    def __init__(self, trades, sigma, sigma_floor: float = 5e-4, sigma_mult: float = 2., ctx=None):
        super().__init__(trades, ctx)
        self.lambda_mult = sigma_mult
        self._sigma = sigma
        self.sigma_floor = sigma_floor

    def _comp_bar_close(self):
        if not (self._device().n == len(self._sigma)):
            raise ValueError("Prices, timestamps, and sigma arrays must have the same length.")
        dev = self._device()
        host = self._sigma.values if isinstance(self._sigma, pd.Series) else self._sigma
        sg = core.device_copy_of(host)           # the sigma a fused Compose(ReturnT, EWMST) left on the device: no second upload
        if isinstance(sg, tuple):
            sg = sg[0]
        if sg is None:
            sg = core.DeviceBuf.upload(dev.ctx, np.ascontiguousarray(host, dtype=np.float64))
        self._dev_index = core.cusum_bar_index(dev, sg, self.sigma_floor, self.lambda_mult)
        # the reference forward-fills NaNs of sigma IN PLACE (logic.py:181-189): mirror it on the caller's array -- only when
        # the device actually filled something (otherwise the 8 B/tick download is skipped)
        if int(dev.ctx._L.fmk_cusum_filled_count(dev.ctx.h)) > 0:
            filled = sg.download(np.float64, len(host))
            if isinstance(host, np.ndarray) and host.dtype == np.float64 and host.flags.writeable:
                host[...] = filled
            else:
                self._sigma = filled
        return self._dev_index.download()

    def get_sigma(self):
        return self._sigma[self.bar_close_indices]


class ImbalanceBarKit(BarBuilderBase):
    """Tick-imbalance bars: ``ImbalanceBarKit(trades, threshold)``.  The reference imports ``_imbalance_bar_indexer`` in
    bar/kit.py:5 but has no kit and the indexer is a stub (logic.py:224-241) -- own semantics, **parity unpinned**: ``b_t`` is
    the ``side`` column when the trades carry one (``use_side``), else the tick rule on the prices."""

    _kind = 0

    def __init__(self, trades, threshold: float, use_side: bool = True, ctx=None):
        super().__init__(trades, ctx)
        self.threshold = threshold
        self.use_side = use_side

    def _comp_bar_close(self):
        self._dev_index = core.imbalance_bar_index(self._device(need_side=self.use_side and self._has_side()),
                                                   self.threshold, self.use_side, self._kind)
        return self._download_index()


class RunBarKit(ImbalanceBarKit):
    """Tick-run bars (stub in the reference, logic.py:244-261) -- own semantics, **parity unpinned**."""

    _kind = 1
