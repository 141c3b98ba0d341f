"""
Trade / footprint containers with the reference's public surface (finmlkit/bar/data_model.py:121-244, :775-1058).

``TradesData`` here is the thin part the hot path needs: the constructor signature, the ``data`` frame (columns
``timestamp, price, amount, id[, side]``, ns ``DatetimeIndex`` named ``datetime``) and the view range.  The ingest
pipeline (``preprocess=True``: id-sort, split-trade merge, tick-rule side) and the HDF5 store are out of scope
 and raise ``NotImplementedError``.  The kits accept the reference's own ``TradesData`` as well: they only
read ``trades.data``.
"""
import datetime as dt
from dataclasses import dataclass
from typing import Optional

import numpy as np
import pandas as pd

from .utils import footprint_to_dataframe


class TradesData:
    def __init__(self, ts, px, qty, id=None, *, is_buyer_maker=None, side=None, dt_index: Optional[pd.DatetimeIndex] = None,
                 timestamp_unit: Optional[str] = None, preprocess: bool = False, proc_res: Optional[str] = None, name=None):
        for nm, a, opt in (("ts", ts, False), ("px", px, False), ("qty", qty, False), ("id", id, True),
                           ("is_buyer_maker", is_buyer_maker, True), ("side", side, True)):
            if a is None and opt:
                continue
            if not isinstance(a, np.ndarray):
                raise TypeError(f"{nm} must be a np.ndarray" if not opt or nm == "id" else f"{nm} must be None or np.ndarray")
        if preprocess:
            raise NotImplementedError("preprocess=True (sort / merge split trades / tick-rule side) is the reference's "
                                      "ingest pipeline and is out of scope here; pass ns timestamps and side=...")
        self._start_date = self._end_date = None
        self._data = pd.DataFrame({'timestamp': ts, 'price': px, 'amount': qty, 'id': id})
        self.is_buyer_maker = is_buyer_maker
        if side is not None:
            self._data['side'] = side
        self._orig_timestamp_unit = timestamp_unit if timestamp_unit else "ns"
        self.name = name
        self.missing_pct = 0
        self.data_ok = None
        self.discontinuities = []
        if dt_index is not None:
            self._data.set_index(dt_index, inplace=True)
        else:
            self._data.set_index(pd.to_datetime(self._data['timestamp'], unit='ns'), inplace=True)
            self._data.index.name = "datetime"

    @property
    def start_date(self):
        return self._start_date

    @property
    def end_date(self):
        return self._end_date

    def set_view_range(self, start, end):
        if isinstance(start, str):
            start = pd.Timestamp(start)
        if isinstance(end, str):
            end = pd.Timestamp(end)
        if start >= end:
            raise ValueError("Start timestamp must be before end timestamp.")
        self._start_date, self._end_date = start, end

    @property
    def data(self) -> pd.DataFrame:
        if self._start_date is None and self._end_date is None:
            return self._data
        return self._data.loc[self._start_date: self._end_date]

    @property
    def orig_timestamp_unit(self) -> str:
        return self._orig_timestamp_unit


def _numba_list(x):
    try:
        from numba.typed import List as NumbaList
    except Exception:  # numba absent: plain lists keep every consumer in this package working
        return list(x)
    return NumbaList(x)


@dataclass
class FootprintData:
    """Same fields and behaviour as the reference dataclass (data_model.py:775-1058)."""
    bar_timestamps: np.ndarray
    price_tick: float
    price_levels: object
    buy_volumes: object
    sell_volumes: object
    buy_ticks: object
    sell_ticks: object
    buy_imbalances: object
    sell_imbalances: object
    cot_price_levels: Optional[np.ndarray] = None
    sell_imbalances_sum: Optional[np.ndarray] = None
    buy_imbalances_sum: Optional[np.ndarray] = None
    imb_max_run_signed: Optional[np.ndarray] = None
    vp_skew: Optional[np.ndarray] = None
    vp_gini: Optional[np.ndarray] = None
    _datetime_index: pd.Series = None

    _RAGGED = ("price_levels", "buy_volumes", "sell_volumes", "buy_ticks", "sell_ticks", "buy_imbalances", "sell_imbalances")
    _PER_BAR = ("cot_price_levels", "sell_imbalances_sum", "buy_imbalances_sum", "imb_max_run_signed", "vp_skew", "vp_gini")

    def __post_init__(self):
        self._datetime_index = pd.to_datetime(self.bar_timestamps, unit='ns')

    def __len__(self) -> int:
        return len(self.bar_timestamps)

    def __getitem__(self, key) -> 'FootprintData':
        if isinstance(key, (slice, int)):
            if isinstance(key, slice) and isinstance(key.start, (str, dt.datetime)) and isinstance(key.stop, (str, dt.datetime)):
                start_idx, end_idx = self._datetime_index.slice_locs(start=key.start, end=key.stop)
                return self[start_idx:end_idx]
            kw = {f: getattr(self, f)[key] for f in self._RAGGED}
            kw.update({f: (getattr(self, f)[key] if getattr(self, f) is not None else None) for f in self._PER_BAR})
            return FootprintData(bar_timestamps=self.bar_timestamps[key], price_tick=self.price_tick, **kw)
        raise TypeError("Invalid argument type. Expected a slice or integer index.")

    @classmethod
    def from_numba(cls, data, price_tick: float) -> 'FootprintData':
        inst = cls(bar_timestamps=np.array(data[0], dtype=np.int64), price_levels=np.array(data[1], dtype=object),
                   price_tick=price_tick, buy_volumes=np.array(data[2], dtype=object), sell_volumes=np.array(data[3], dtype=object),
                   buy_ticks=np.array(data[4], dtype=object), sell_ticks=np.array(data[5], dtype=object),
                   buy_imbalances=np.array(data[6], dtype=object), sell_imbalances=np.array(data[7], dtype=object),
                   buy_imbalances_sum=np.array(data[8], dtype=np.uint16), sell_imbalances_sum=np.array(data[9], dtype=np.uint16),
                   cot_price_levels=np.array(data[10], dtype=np.int32), imb_max_run_signed=np.array(data[11], dtype=np.int16),
                   vp_skew=np.array(data[12], dtype=np.float64), vp_gini=np.array(data[13], dtype=np.float64))
        if not inst.is_valid():
            raise ValueError("Inconsistent data length in the FootprintData container!")
        return inst

    @classmethod
    def from_csr(cls, bar_timestamps, price_tick, csr) -> 'FootprintData':
        """Build from the device CSR tuple (core.bar_footprints_csr); ragged members become lists of array views."""
        off = csr[0]
        ragged = [[x[off[i]:off[i + 1]] for i in range(len(off) - 1)] for x in csr[1:8]]
        return cls(bar_timestamps=np.asarray(bar_timestamps, dtype=np.int64), price_tick=price_tick,
                   price_levels=ragged[0], buy_volumes=ragged[1], sell_volumes=ragged[2], buy_ticks=ragged[3],
                   sell_ticks=ragged[4], buy_imbalances=ragged[5], sell_imbalances=ragged[6],
                   buy_imbalances_sum=csr[8], sell_imbalances_sum=csr[9], cot_price_levels=csr[10],
                   imb_max_run_signed=csr[11], vp_skew=csr[12], vp_gini=csr[13])

    def get_df(self):
        return footprint_to_dataframe(self.bar_timestamps, self.price_levels, self.buy_volumes, self.sell_volumes,
                                      self.buy_ticks, self.sell_ticks, self.buy_imbalances, self.sell_imbalances, self.price_tick)

    def cast_to_numba_list(self):
        for f in self._RAGGED:
            setattr(self, f, _numba_list(getattr(self, f)))

    def cast_to_numpy(self):
        for f in self._RAGGED:
            xs = getattr(self, f)
            arr = np.empty(len(xs), dtype=object)
            for i, x in enumerate(xs):
                arr[i] = x
            setattr(self, f, arr)

    def memory_usage(self):
        total = 0
        for f in self._RAGGED:
            total += sum(np.asarray(x).nbytes for x in getattr(self, f))
        for f in self._PER_BAR + ("bar_timestamps",):
            a = getattr(self, f)
            if a is not None:
                total += np.asarray(a).nbytes
        return total / (1024 ** 2)

    def is_valid(self) -> bool:
        n = len(self.bar_timestamps)
        return all(len(getattr(self, f)) == n for f in self._RAGGED)
