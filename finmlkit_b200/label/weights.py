"""label/weights.py of the reference: ``average_uniqueness`` and ``return_attribution`` run on the GPU (csrc/weights.cu,
O(N + E) instead of O(sum of label path lengths)); ``time_decay`` and ``class_balance_weights`` are O(n_events) host
arithmetic in the reference's evaluation order (weights.py:106-199)."""
import numpy as np

from .. import core


def average_uniqueness(timestamps, event_idxs, touch_idxs, ctx=None):
    """weights.py:7-49 -> ``(weights f64[E], concurrency int16[n])``.  Only ``len(timestamps)`` is used, as in the
    reference.  Indices must lie in ``[0, n)``."""
    if len(event_idxs) != len(touch_idxs):
        raise ValueError("Timestamps and lookahead indices must have the same length.")
    return core.average_uniqueness_dev(len(timestamps), event_idxs, touch_idxs, ctx=ctx)


def return_attribution(event_idxs, touch_idxs, close, concurrency, normalize, ctx=None):
    """weights.py:52-103."""
    return core.return_attribution_dev(event_idxs, touch_idxs, close, concurrency, normalize, ctx=ctx)


def time_decay(avg_uniqueness, last_weight):
    """weights.py:106-143 (sequential cumsum, like ``np.cumsum`` under Numba)."""
    if not -1.0 <= last_weight <= 1.0:
        raise ValueError("last_weight must lie in [-1, 1]")
    cum = np.cumsum(np.asarray(avg_uniqueness, dtype=np.float64))
    if cum[-1] == 0.0:
        raise ValueError("The sum of all average uniqueness weights must be grater than 0.")
    if last_weight >= 0.0:
        slope = (1. - last_weight) / cum[-1]
    else:
        slope = 1. / ((last_weight + 1.) * cum[-1])
    const = 1. - slope * cum[-1]
    weights = const + slope * cum
    if last_weight < 0.0:
        weights = np.maximum(weights, 0.0)
    return weights


def class_balance_weights(labels, base_w):
    """weights.py:147-199 -> ``(unique_labels, class_weights, sum_w_class, final_weights)``."""
    labels = np.asarray(labels)
    base_w = np.asarray(base_w, dtype=np.float64)
    unique_labels = np.unique(labels)
    n_classes = len(unique_labels)
    label_idx = np.searchsorted(unique_labels, labels)
    sum_w_class = np.zeros(n_classes, dtype=np.float64)
    np.add.at(sum_w_class, label_idx, base_w)            # sequential accumulation in sample order, like the loop
    total = 0.0
    for c in range(n_classes):                           # np.sum of a handful of classes
        total += sum_w_class[c]
    class_weights = np.zeros(n_classes, dtype=np.float64)
    for c in range(n_classes):
        class_weights[c] = total / (n_classes * sum_w_class[c]) if sum_w_class[c] > 0. else 0.0
    final_weights = base_w * class_weights[label_idx]
    return unique_labels, class_weights, sum_w_class, final_weights
