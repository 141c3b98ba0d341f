"""``TBMLabel`` with the reference's constructor, properties and outputs (label/kit.py:12-322)."""
import numpy as np
import pandas as pd

from .tbm import triple_barrier


class TBMLabel:
    def __init__(self, features: pd.DataFrame, target_ret_col: str, min_ret: float, horizontal_barriers: tuple,
                 vertical_barrier: pd.Timedelta, min_close_time: pd.Timedelta = pd.Timedelta(seconds=1), is_meta: bool = False):
        if target_ret_col not in features.columns:
            raise ValueError(f"Target column '{target_ret_col}' not found in features DataFrame.")
        if not isinstance(features.index, pd.DatetimeIndex):
            raise ValueError("Features index must be a DatetimeIndex.")
        if not isinstance(horizontal_barriers, tuple) or len(horizontal_barriers) != 2:
            raise ValueError("Horizontal barriers must be a tuple of two floats (bottom, top).")
        if min_ret < 0.:
            raise ValueError("Minimum return must be non-negative.")
        if is_meta:
            if 'side' not in features.columns:
                raise ValueError("For meta labeling, 'side' column must be present in features DataFrame.")
            if not pd.api.types.is_integer_dtype(features['side']):
                raise ValueError("The 'side' column must be of integer type (e.g., -1, 0, 1).")
        self._orig_features = self._preprocess_features(features, target_ret_col, min_ret, horizontal_barriers)
        self._features = self._orig_features
        self.target_ret_col = target_ret_col
        self.min_ret = min_ret
        self.horizontal_barriers = horizontal_barriers
        self.vertical_barrier = vertical_barrier.total_seconds()
        self.min_close_time_sec = min_close_time.total_seconds()
        self.is_meta = is_meta
        self._out = None

    @staticmethod
    def _preprocess_features(x, target_ret_col, min_ret, horizontal_barriers):
        first_valid = [x[c].first_valid_index() for c in x.columns if x[c].first_valid_index() is not None]
        if not first_valid:
            raise ValueError("All columns contain only NaN values.")
        x = x.loc[max(first_valid):]
        x = x[x[target_ret_col].abs() * np.max(horizontal_barriers) >= min_ret]
        if x.empty:
            raise ValueError("No valid events found after filtering by minimum return and removing leading NaNs.")
        if x[target_ret_col].isna().any():
            raise ValueError(f"Target return column '{target_ret_col}' contains NaN values. Please ensure it is fully populated.")
        return x

    @property
    def event_count(self):
        return len(self._features)

    @property
    def features(self):
        return self._features

    @property
    def target_returns(self):
        return self._features[self.target_ret_col]

    @property
    def labels(self):
        if self._out is None:
            raise ValueError("Labels have not been computed yet. Call `compute_labels()` first.")
        return self._out['labels']

    @property
    def event_returns(self):
        if self._out is None or 'returns' not in self._out.columns:
            raise ValueError("Log returns have not been computed yet. Call `compute_labels()` first.")
        return self._out['returns']

    @property
    def full_output(self):
        if self._out is None:
            raise ValueError("Labels have not been computed yet. Call `compute_labels()` and `compute_weights` first.")
        return self._out

    def _drop_trailing_events(self, trades):
        last = pd.Timestamp(trades.data.timestamp.values[-1], unit='ns')
        return self._orig_features[self._orig_features.index + pd.Timedelta(self.vertical_barrier, unit="s") <= last]

    def compute_labels(self, trades):
        """label/kit.py:272-313 -> (features, out) with columns touch_time, event_idx, touch_idx, labels, returns,
        vertical_touch_weights."""
        if not hasattr(trades, "data"):
            raise ValueError("Trades must be an instance of TradesData.")
        self._features = self._drop_trailing_events(trades)
        ts = trades.data.timestamp.values.astype(np.int64)
        if "event_idx" in self._features.columns:
            event_idx = self._features.event_idx.values
        else:
            event_idx = np.searchsorted(ts, self._features.index.as_unit("ns").asi8)
        from .. import core
        # the trades frame's ONE device copy (shared with the bar kits and the sigma transforms): no second upload
        dev = core.device_trades_for(trades.data, need_ts=True)
        labels, touch_idx, rets, ratios = triple_barrier(
            timestamps=ts, close=trades.data.price.values, event_idxs=event_idx, targets=self.target_returns.values,
            horizontal_barriers=self.horizontal_barriers, vertical_barrier=self.vertical_barrier,
            min_close_time_sec=self.min_close_time_sec,
            side=self.features['side'].values.astype(np.int8) if self.is_meta else None, min_ret=self.min_ret, _dev_trades=dev)
        self._out = pd.DataFrame({'touch_time': pd.to_datetime(ts[touch_idx]), 'event_idx': event_idx, 'touch_idx': touch_idx,
                                  'labels': labels, 'returns': rets, 'vertical_touch_weights': ratios}, index=self.features.index)
        return self.features, self.full_output

    def compute_weights(self, trades, normalized: bool = False) -> pd.DataFrame:
        """label/kit.py:315-322."""
        return SampleWeights.compute_info_weights(trades, self._out, normalized)


class SampleWeights:
    """label/kit.py:325-477: average uniqueness + return attribution (GPU, one pass over the price column), then the
    O(n_events) time-decay / class-balance combination."""

    @staticmethod
    def compute_info_weights(trades, labels: pd.DataFrame, normalize: bool = False) -> pd.DataFrame:
        if not hasattr(trades, "data"):
            raise ValueError("Trades must be an instance of TradesData.")
        if not isinstance(labels, pd.DataFrame):
            raise ValueError("Events must be a pandas DataFrame.")
        if 'event_idx' not in labels.columns or 'touch_idx' not in labels.columns:
            raise ValueError("Events DataFrame must contain 'event_idx' and 'touch_idxs' columns.")
        from .. import core
        tr = core.device_trades_for(trades.data)               # the shared device copy of the frame; only price is read
        avg_u, info_w = core.sample_weights_dev(tr, labels.event_idx.values, labels.touch_idx.values, normalize=normalize)
        out_df = pd.DataFrame({'avg_uniqueness': avg_u}, index=labels.index)
        out_df["return_attribution"] = info_w
        return out_df

    @staticmethod
    def compute_final_weights(avg_uniqueness: pd.Series, time_decay_intercept: float = 1., return_attribution: pd.Series = None,
                              vertical_touch_weights: pd.Series = None, labels: pd.Series = None) -> pd.DataFrame:
        from .weights import time_decay, class_balance_weights
        if not isinstance(avg_uniqueness, pd.Series):
            raise ValueError("avg_uniqueness must be a pandas Series.")
        if not isinstance(time_decay_intercept, (int, float)):
            raise ValueError("time_decay_intercept must be a numeric value.")
        if not -1.0 <= time_decay_intercept <= 1.0:
            raise ValueError("time_decay_intercept must lie in [-1, 1]")
        for name, s in (("return_attribution", return_attribution), ("vertical_touch_weights", vertical_touch_weights), ("labels", labels)):
            if s is not None and not isinstance(s, pd.Series):
                raise ValueError(f"{name} must be a pandas Series.")
        for name, s in (("return_attribution", return_attribution), ("vertical_touch_weights", vertical_touch_weights), ("labels", labels)):
            if s is not None and not avg_uniqueness.index.equals(s.index):
                raise ValueError(f"avg_uniqueness and {name} must have the same index.")
        n_events = len(avg_uniqueness)
        time_decay_weights = time_decay(avg_uniqueness.values, time_decay_intercept)
        out_df = pd.DataFrame({'time_decay_weights': time_decay_weights}, index=avg_uniqueness.index)
        if return_attribution is not None:
            if return_attribution.sum() <= 0:
                raise ValueError("Return attribution sum is zero or negative, cannot normalize.")
            ra = return_attribution.values * n_events / return_attribution.sum()
            out_df["return_attribution"] = ra
            combined = time_decay_weights * ra
        else:
            combined = time_decay_weights * avg_uniqueness.values
        if vertical_touch_weights is not None:
            out_df["vertical_touch_weights"] = vertical_touch_weights.values
            combined = combined * vertical_touch_weights.values
        mean_combined = combined.mean()
        if mean_combined <= 0:
            raise ValueError("Mean of combined weights is zero or negative, cannot normalize.")
        base_weights = combined / mean_combined
        if labels is not None:
            _, _, _, final_weights = class_balance_weights(labels.values, base_weights)
        else:
            final_weights = base_weights
        out_df["weights"] = final_weights
        return out_df
