"""Known-answer vectors re-hosted from the reference's own unit tests (hand-computed expectations, cited per case).
Used twice: against the CPU oracle (tests/test_oracle_vectors.py) and against the CUDA path through the mirrored
Python API (tests/test_gpu_vectors.py)."""
import numpy as np

F, I = np.float64, np.int64

# comp_bar_ohlcv -- reference tests/bars/test_comp_ohlcv.py:8-135
OHLCV = [
    dict(name="three_and_two_trades", p=[10., 11., 12., 13., 14., 15.], v=[100., 200., 150., 100., 50., 25.], idx=[0, 3, 5],
         open=[11., 14.], high=[13., 15.], low=[11., 14.], close=[13., 15.], volume=[450., 75.],
         vwap=[(11 * 200 + 12 * 150 + 13 * 100) / 450., (14 * 50 + 15 * 25) / 75.], trades=[3, 2], median=[150., 37.5]),
    dict(name="single_trade_per_bar", p=[10., 12., 14.], v=[100., 200., 300.], idx=[0, 1, 2],
         open=[12., 14.], high=[12., 14.], low=[12., 14.], close=[12., 14.], volume=[200., 300.], vwap=[12., 14.],
         trades=[1, 1], median=[200., 300.]),
    dict(name="zero_volume", p=[10., 12.], v=[100., 0.], idx=[0, 1],
         open=[12.], high=[12.], low=[12.], close=[12.], volume=[0.], vwap=[0.], trades=[1], median=[0.]),
    dict(name="empty_bar_in_the_middle", p=[10., 12., 14.], v=[100., 200., 300.], idx=[0, 1, 1, 2],
         open=[12., 12., 14.], high=[12., 12., 14.], low=[12., 12., 14.], close=[12., 12., 14.], volume=[200., 0., 300.],
         vwap=[12., 0., 14.], trades=[1, 0, 1], median=[200., 0., 300.]),
]

# _time_bar_indexer clock -- reference tests/bars/test_time_bar_indexer.py:9-80
S = 1_000_000_000
TIME_CLOCK = [
    dict(ts=[999_999_999, 1 * S, 2 * S, 3 * S, 4 * S, 5 * S, 5_999_999_999, 6_100_000_000, 7 * S], interval=2.0,
         clock=[0, 2 * S, 4 * S, 6 * S, 8 * S, 10 * S]),
    dict(ts=[1 * S, 1_500_000_000, 2 * S, 2_500_000_000, 3 * S, 4 * S], interval=2.0, clock=[0, 2 * S, 4 * S, 6 * S]),
    dict(ts=[1 * S, 2 * S, 5 * S, 6 * S], interval=2.0, clock=[0, 2 * S, 4 * S, 6 * S, 8 * S]),
]

# comp_bar_directional_features -- reference tests/bars/test_comp_bar_directional_features.py:33-75
DIRECTIONAL = [
    dict(p=[100., 101., 102.], v=[10., 15., 20.], side=[0, 1, 1], idx=[0, 2],
         out=[[2], [0], [35.], [0.], [3555.], [0.], [0.5], [1.], [1], [2], [15.], [35.], [1515.], [3555.]]),
    dict(p=[100., 99., 100., 101.], v=[10., 20., 30., 40.], side=[0, -1, 1, 1], idx=[0, 3],
         out=[[2], [1], [70.], [20.], [7040.], [1980.], [0.6666667], [1.], None, None, None, None, None, None]),
]

# comp_bar_footprints -- reference tests/bars/test_comp_bar_footprints.py:8-60
FOOTPRINT = dict(p=[100.0, 100.5, 101.0, 100.5, 100.0], a=[1.0, 2.0, 1.5, 1.0, 2.0], idx=[0, 3, 4], side=[0, 1, 1, -1, -1],
                 tick=0.5, lows=[100.0, 100.0], highs=[101.0, 100.0], factor=1.5,
                 levels0=[200, 201, 202], buy0=[0.0, 2.0, 1.5], sell0=[0.0, 1.0, 0.0], bt0=[0, 1, 1], st0=[0, 1, 0])

# comp_price_tick_size -- reference tests/bars/test_utils.py:30-66
TICK_SIZE = [([100.0, 100.5, 101.0, 101.5, 102.0], 0.5), ([1.0], 0.0), ([5.0, 5.0, 5.0], 0.0),
             ([30000.1, 30000.3, 30000.2, 29999.9], 0.1), ([0.01, 0.03, 0.02], 0.01)]

# triple_barrier -- reference tests/labels/test_triple_barrier.py:195-246
TBM = [
    dict(close=[100, 110, 120, 130, 140], tgt=0.05, vert=10.0, label=1, ratio=1.0, touch_lt=4),
    dict(close=[100, 90, 80, 70, 60], tgt=0.05, vert=10.0, label=-1, ratio=1.0, touch_lt=4),
]
TBM_TIMEOUT = dict(close=[100, 100.5, 101, 100.5, 100, 100.2, 99.8, 100.1, 99.9, 100.3], tgt=0.1, vert=5.0, touch=5)

# triple_barrier argument errors -- reference label/tbm.py:45-59, tests/labels/test_triple_barrier.py:35-69
TBM_ERRORS = [
    (dict(vert=0.0), "The vertical barrier must be greater than zero."),
    (dict(min_ret=-1.0), "The minimum return must be non-negative."),
    (dict(short_close=True), "The lengths of timestamps and close must match."),
    (dict(short_targets=True), "The lengths of event_idxs and targets must match."),
    (dict(no_events=True), "The event_idxs array must not be empty."),
    (dict(bad_side=True), "The length of event_idxs must match the length of side."),
]


def tbm_error_args(kind):
    ts = np.arange(5, dtype=I) * S
    close = np.array([100., 101., 102., 103., 104.])
    ev, tg = np.array([0, 1], I), np.array([0.01, 0.01])
    side = None
    vert, min_ret = kind.get("vert", 10.0), kind.get("min_ret", 0.0)
    if kind.get("short_close"):
        close = close[:-1]
    if kind.get("short_targets"):
        tg = tg[:-1]
    if kind.get("no_events"):
        ev, tg = np.zeros(0, I), np.zeros(0)
    if kind.get("bad_side"):
        side = np.ones(1, np.int8)
    return ts, close, ev, tg, (1.0, 1.0), vert, 0.0, side, min_ret


# average_uniqueness -- reference tests/labels/test_average_uniqueness.py:7-62 (hand-computed concurrency / weights)
AVG_UNIQUENESS = [
    dict(n=10, ev=[0, 4, 8], touch=[2, 6, 9], w=[1.0, 1.0, 1.0], conc=[1, 1, 1, 0, 1, 1, 1, 0, 1, 1]),
    dict(n=5, ev=[0, 0, 0], touch=[4, 4, 4], w=[1 / 3, 1 / 3, 1 / 3], conc=[3, 3, 3, 3, 3]),
    dict(n=8, ev=[0, 2, 4], touch=[3, 5, 7], w=[0.75, 0.5, 0.75], conc=[1, 1, 2, 2, 2, 2, 1, 1]),
]
# return_attribution -- reference tests/labels/test_return_attribution.py:26-75
RETURN_ATTRIBUTION = dict(close=[100., 102., 104., 106.], ev=[0], touch=[3], conc=[1, 1, 1, 1])
RETURN_ATTRIBUTION_ZERO = dict(close=[100., 100., 100., 100.], ev=[0, 2], touch=[1, 3], conc=[1, 1, 1, 1])
# time_decay -- reference tests/labels/test_time_decay.py:22-47
TIME_DECAY = [([0.5, 0.5, 0.5, 0.5], 1.0, [1.0, 1.0, 1.0, 1.0]), ([0.5, 0.5, 0.5, 0.5], 0.4, [0.55, 0.7, 0.85, 1.0])]

# class_balance_weights -- reference tests/labels/test_class_balace_weights.py:10-75 (hand-computed)
CLASS_BALANCE = [
    dict(labels=[1, -1, 1, 0, -1, 1], w=[1., 1., 1., 1., 1., 1.], uniq=[-1, 0, 1], sums=[2., 1., 3.], cw=[1.0, 2.0, 6 / 9]),
    dict(labels=[1, -1, 1, 0, -1, 1], w=[1., 2., 1., 3., 2., 1.], uniq=[-1, 0, 1], sums=[4., 3., 3.], cw=[10 / 12, 10 / 9, 10 / 9]),
]
# calc_volume_percentage_above_poc through volume_profile_rolling on a single aggregated window -- reference
# tests/features/test_volume_profile_rolling.py:13-50: five levels 100..104, POC at the max-volume level
VP_ABOVE_POC = [
    dict(vol=[10., 5., 20., 5., 10.], poc=102, pct=0.3),
    dict(vol=[0., 0., 5., 10., 15.], poc=104, pct=0.0),
    dict(vol=[10., 5., 0., 0., 0.], poc=100, pct=5. / 15.),
]
