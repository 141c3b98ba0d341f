"""sampling/filters.py of the reference: the symmetric CUSUM event filter on the GPU (csrc/index_cusum.cu, the same
chunk-chain fix-point as the CUSUM bar indexer, with the filter's strict inequalities and s- / s+ test order)."""
import numpy as np

from .. import core


def cusum_filter(raw_time_series, threshold, ctx=None):
    """filters.py:6-70 -> int64 indices of the events (positions in ``raw_time_series``)."""
    if len(raw_time_series) <= 1:
        raise ValueError("Input time series must have at least 2 elements.")
    threshold = np.atleast_1d(np.asarray(threshold, dtype=np.float64))
    if len(threshold) != 1 and len(threshold) != len(raw_time_series):
        raise ValueError("Threshold array must either contain 1 const. element or len(raw_time_series) elements.")
    return core.cusum_filter_dev(raw_time_series, threshold, ctx=ctx)
