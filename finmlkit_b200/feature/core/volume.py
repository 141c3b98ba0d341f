"""feature/core/volume.py:572-641 (``comp_flow_acceleration``, ``vpin``) on the GPU."""
from ... import core


def vpin(volume_buy, volume_sell, window, ctx=None):
    return core.vpin_series(volume_buy, volume_sell, window, ctx=ctx)


def comp_flow_acceleration(volumes, window, recent_periods, ctx=None):
    return core.flow_acceleration_series(volumes, window, recent_periods, ctx=ctx)
