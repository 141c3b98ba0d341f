"""GPU: round-2 additions -- the device-resident bar frame, float32 amount ingest, device gathers, index validation, the
lagged-returns bracket edge (span == staging size), NaN volumes in flow acceleration, and libfmk's NCCL communicator."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from helpers import assert_exact, assert_f64, check_directional, check_footprint_csr, check_ohlcv, check_trade_size

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stream(n=200_000, seed=11):
    from finmlkit_b200.synth import synth_trades
    return synth_trades(n, seed=seed)


def test_lagged_returns_bracket_span_equals_stage(ctx):
    """a 1024-tick block whose bracket [b0, b1] spans exactly 4096 staged timestamps -- the block's last tick
    has its insertion point at b1 = b0 + 4096, one more than the halving steps 2048..1 can reach."""
    from finmlkit_b200 import core
    dense = np.arange(20480, dtype=np.int64)                                   # 1 ns apart
    blk = 30000 + np.round(np.linspace(0, 4096, 1024)).astype(np.int64)       # one aligned block spanning 4096 ns
    tail = blk[-1] + 5 + np.arange(700, dtype=np.int64) * 3
    ts = np.concatenate([dense, blk, tail])
    px = 100.0 + np.cumsum(np.random.default_rng(3).normal(0, 0.01, len(ts)))
    w_sec = 2e-5
    w = w_sec * 1e9
    tf = ts.astype(np.float64)
    b0 = np.searchsorted(tf, tf[20480] - w, "right")
    b1 = np.searchsorted(tf, tf[20480 + 1023] - w, "right")
    assert b1 - b0 == 4096, (b0, b1)                                           # the case the test is about
    for is_log in (True, False):
        got = core.lagged_returns(ts, px, w_sec, is_log, ctx=ctx)
        ref = oracle.comp_lagged_returns(ts, px, w_sec, is_log)
        assert_f64(got, ref, f"lagged returns span 4096 log={is_log}", rtol=1e-12, atol=1e-15)
    # and the neighbouring spans (4095, 4097) for good measure
    for extra in (-1, 1):
        blk2 = 30000 + np.round(np.linspace(0, 4096 + extra, 1024)).astype(np.int64)
        ts2 = np.concatenate([dense, blk2, blk2[-1] + 5 + np.arange(700, dtype=np.int64) * 3])
        assert_f64(core.lagged_returns(ts2, px, w_sec, True, ctx=ctx), oracle.comp_lagged_returns(ts2, px, w_sec, True),
                   f"span {4096 + extra}", rtol=1e-12, atol=1e-15)


def test_flow_acceleration_nan_poisons_later_bars(ctx):
    """comp_flow_acceleration (volume.py:572-607) has no NaN handling -- a NaN volume makes every later output NaN."""
    from finmlkit_b200.feature.core.volume import comp_flow_acceleration
    v = np.abs(np.random.default_rng(5).normal(10, 2, 5000))
    v[1234] = np.nan
    got = comp_flow_acceleration(v, 20, 5, ctx=ctx)
    ref = oracle.comp_flow_acceleration(v, 20, 5)
    assert np.isnan(got[1234:]).all() and np.isfinite(got[19:1234]).all()
    assert_f64(got, ref, "flow acceleration with a NaN volume")


def test_caller_indices_are_validated(ctx):
    """comp_bar_* with an index >= n, < -1 or a decreasing pair must raise, not read out of bounds."""
    from finmlkit_b200.bar.base import comp_bar_directional_features, comp_bar_ohlcv
    ts, px, qty, side = _stream(5000)
    for bad in ([0, 100, 5000], [-2, 10, 20], [0, 300, 200, 400]):
        with pytest.raises(ValueError):
            comp_bar_ohlcv(px, qty, np.array(bad, np.int64), ctx=ctx)
        with pytest.raises(ValueError):
            comp_bar_directional_features(px, qty, np.array(bad, np.int64), side, ctx=ctx)
    # -1 (time bars) and repeated indices (empty bars) are legal
    o = comp_bar_ohlcv(px, qty, np.array([-1, 10, 10, 4999], np.int64), ctx=ctx)
    check_ohlcv(o, oracle.comp_bar_ohlcv(px, qty, np.array([-1, 10, 10, 4999], np.int64)), "legal indices")
    # the context is still healthy afterwards
    ctx.sync()


def test_float32_amount_upload(ctx):
    """TradesData after the reference's split-trade merge holds float32 amounts (data_model.py:326-344): they cross PCIe as
    float32 and every result equals the float64 path on the widened values."""
    from finmlkit_b200 import core
    ts, px, qty, side = _stream(120_000)
    q32 = qty.astype(np.float32)
    tr = core.DeviceTrades.upload(ts, px, q32, side, ctx=ctx)
    back = tr.download()
    assert back[2].dtype == np.float64 and np.array_equal(back[2], q32.astype(np.float64))
    q64 = q32.astype(np.float64)
    ix = core.dollar_bar_index(tr, 2e5)
    ref = oracle.dollar_bar_indexer(px, q64, 2e5)
    assert_exact(ix.download()[1], ref, "dollar idx on float32 amounts")
    check_ohlcv(core.bar_ohlcv(tr, ix), oracle.comp_bar_ohlcv(px, q64, ref), "ohlcv on float32 amounts")


def test_buf_gather(ctx):
    from finmlkit_b200 import core
    x = np.random.default_rng(1).normal(size=10_000)
    b = core.DeviceBuf.upload(ctx, x)
    idx = np.array([0, 9999, 17, 17, 4242, -1], np.int64)
    assert np.array_equal(b.gather(idx), x[idx])
    assert len(b.gather(np.zeros(0, np.int64))) == 0


@pytest.mark.parametrize("kind", ["dollar", "volume", "time"])
def test_frame_equals_the_individual_builders(ctx, kind):
    """fmk_bar_features_device = build_ohlcv + build_directional_features + build_trade_size_features + build_footprints in
    one call with every column kept on the device: bit-identical to the one-at-a-time C-ABI calls, and parity with the oracle."""
    from finmlkit_b200 import core
    ts, px, qty, side = _stream(300_000, seed=21)
    side[::97] = 0                                                # some side-0 ticks
    tr = core.DeviceTrades.upload(ts, px, qty, side, ctx=ctx)
    ix = {"dollar": lambda: core.dollar_bar_index(tr, 1.5e5), "volume": lambda: core.volume_bar_index(tr, 4.0),
          "time": lambda: core.time_bar_index(tr, 60.0)}[kind]()
    cts, cidx = ix.download()
    fr = core.bar_features_device(tr, ix, core.F_ALL, theta=None, theta_mult=5.0, price_tick_size=0.1, imbalance_factor=3.0)
    c = fr.download()
    nb = len(cidx) - 1
    assert fr.n_bars == nb and np.array_equal(c["close_idx"], cidx[1:]) and np.array_equal(c["close_ts"], cts[1:])
    o = core.bar_ohlcv(tr, ix)
    for k, name in enumerate(["open", "high", "low", "close", "volume", "vwap", "trades", "median_trade_size"]):
        assert_exact(c[name], o[k], f"frame.{name}")
    d = core.bar_directional(tr, ix)
    names = ["ticks_buy", "ticks_sell", "volume_buy", "volume_sell", "dollars_buy", "dollars_sell", "mean_spread", "max_spread",
             "cum_ticks_min", "cum_ticks_max", "cum_volume_min", "cum_volume_max", "cum_dollars_min", "cum_dollars_max"]
    for k, name in enumerate(names):
        assert_exact(c[name], d[k], f"frame.{name}")
    t = core.bar_trade_size(tr, ix, o[7], 5.0)
    for k, name in enumerate(["mean_size_rel", "size_95_rel", "pct_block", "size_gini"]):
        assert_exact(c[name], t[k], f"frame.{name}")
    f = core.bar_footprints_csr(tr, ix, 0.1, o[2], o[1], 3.0)
    fnames = ["fp_level_offsets", "fp_price_levels", "fp_buy_vol", "fp_sell_vol", "fp_buy_ticks", "fp_sell_ticks", "fp_buy_imb",
              "fp_sell_imb", "fp_buy_imb_sum", "fp_sell_imb_sum", "fp_cot", "fp_run_signed", "fp_vp_skew", "fp_vp_gini"]
    for k, name in enumerate(fnames):
        assert_exact(c[name], f[k], f"frame.{name}")
    # and against the oracle (the reference's own evaluation order)
    check_ohlcv(o, oracle.comp_bar_ohlcv(px, qty, cidx), kind)
    check_directional(d, oracle.comp_bar_directional_features(px, qty, cidx, side), kind)
    check_trade_size(t, oracle.comp_bar_trade_size_features(qty, o[7], cidx, 5.0), kind)
    fo = oracle.comp_bar_footprints_csr(px, qty, cidx, side, 0.1, o[2], o[1], 3.0)
    check_footprint_csr(f, fo[0], list(fo[1:]), float(np.max(np.abs(fo[1]))), kind)
    # partial frames: absent columns are absent, present ones unchanged
    fr2 = core.bar_features_device(tr, ix, core.F_OHLCV)
    c2 = fr2.download()
    assert "median_trade_size" not in c2 and "ticks_buy" not in c2 and fr2.level_bytes == 0
    assert_exact(c2["high"], o[1], "partial frame high")
    with pytest.raises(ValueError):
        core.bar_features_device(tr, ix, core.F_FOOTPRINT)      # needs OHLCV for the lows / highs


def test_trade_size_tolerance_is_one_ulp(ctx):
    """the trade-size features are float64 computations cast to float32 -> within ONE float32 ulp."""
    from finmlkit_b200 import core
    from helpers import assert_f32_ulp
    ts, px, qty, side = _stream(400_000, seed=33)
    tr = core.DeviceTrades.upload(ts, px, qty, side, ctx=ctx)
    for T in (2e4, 2e5, 3e6):
        ix = core.dollar_bar_index(tr, T)
        cidx = ix.download()[1]
        theta = oracle.comp_bar_ohlcv(px, qty, cidx)[7]
        got = core.bar_trade_size(tr, ix, theta, 5.0)
        ref = oracle.comp_bar_trade_size_features(qty, theta, cidx, 5.0)
        for k in range(4):
            assert_f32_ulp(got[k], ref[k], f"T={T} tsize[{k}]", ulps=1)


def _frame_bytes(core, ctx, fr):
    bar = np.zeros((fr.bar_bytes + 15) // 16 * 16, np.uint8)
    lvl = np.zeros((fr.level_bytes + 15) // 16 * 16, np.uint8)
    ctx.check(ctx._L.fmk_frame_download(ctx.h, fr.h, bar.ctypes.data_as(__import__("ctypes").c_void_p),
                                        lvl.ctypes.data_as(__import__("ctypes").c_void_p) if fr.level_bytes else None))
    return np.concatenate([bar, lvl]) if fr.level_bytes else bar


def test_comm_single_rank_gather(ctx):
    """libfmk's communicator with world = 1 (NCCL loaded with dlopen, no torch): the gathered frame on rank 0 is byte for
    byte the packed frame, across pipelined steps of changing size."""
    import ctypes as C
    from finmlkit_b200 import core
    from finmlkit_b200.parallel import Comm
    L = ctx._L
    uid = C.create_string_buffer(128)
    assert L.fmk_comm_unique_id(uid) == 0 and L.fmk_comm_nccl_version() > 20000
    comm = Comm(ctx, 0, 1, uid.raw, max_ctas=4)
    try:
        comm.barrier()
        assert np.array_equal(comm.allreduce([1.5, -2.0], "max"), [1.5, -2.0])
        ts, px, qty, side = _stream(150_000, seed=5)
        tr = core.DeviceTrades.upload(ts, px, qty, side, ctx=ctx)
        expect = None
        for T in (1e5, 3e5, 5e4):
            ix = core.dollar_bar_index(tr, T)
            fr = core.bar_features_device(tr, ix, core.F_ALL, price_tick_size=0.1)
            expect = _frame_bytes(core, ctx, fr)
            comm.gather_submit(fr.segments(), dst=0)
            del fr, ix
        comm.gather_finish()
        ctx.sync()
        got = comm.gathered_frame(0)
        assert comm.gathered_bytes() == [len(expect)]
        assert np.array_equal(got, expect)
        # the gathered bytes describe themselves (512-byte header): the receiver rebuilds every column without side information
        cols = core.frame_views_from_bytes(got)
        ix = core.dollar_bar_index(tr, 5e4)
        ref = core.bar_features_device(tr, ix, core.F_ALL, price_tick_size=0.1).download()
        assert set(cols) == set(ref)
        for k in ref:
            assert_exact(cols[k], ref[k], f"gathered frame column {k}")
    finally:
        comm.destroy()


@pytest.mark.parametrize("p2p,reset", [("1", ""), ("1", "1"), ("0", "")])
def test_comm_two_ranks_gather_equals_each_ranks_frame(tmp_path, p2p, reset):
    """world = 2 on two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`): rank 0 receives exactly the bytes rank 1
    packed, for frames of different sizes, with the transfer of step k overlapping step k+1 -- through the peer-to-peer push
    (CUDA IPC mapping of rank 0's receive buffer, with and without a collective reset in the middle) and through the
    ncclSend / ncclRecv fallback."""
    from finmlkit_b200 import _lib
    if _lib.lib().fmk_device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    env.update({"WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29617", "FMK_STORE_DIR": str(tmp_path),
                "FMK_COMM_TEST_DIR": str(tmp_path), "FMK_COMM_P2P": p2p, "FMK_COMM_TEST_RESET": reset})
    procs = []
    for r in range(2):
        e = dict(env)
        e.update({"RANK": str(r), "LOCAL_RANK": str(r)})
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_comm_worker.py")], env=e,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    try:
        outs = [p.communicate(timeout=150)[0] for p in procs]
    except subprocess.TimeoutExpired:
        for p in procs:
            p.kill()
        pytest.fail("two-rank communicator test hung: " + " | ".join((p.communicate()[0] or "")[-1500:] for p in procs))
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r}:\n{outs[r][-3000:]}"
    assert "GATHER_OK" in outs[0]
    assert ("p2p-ipc-copy-engine" if p2p == "1" else "nccl-send-recv") in outs[0], outs[0][-2000:]


@pytest.mark.parametrize("kind,thr", [(0, 7.0), (0, 60.0), (0, 2.5), (1, 40.0), (1, 900.0)])
def test_imbalance_and_run_bars_match_the_own_oracle(ctx, kind, thr, monkeypatch):
    """a6 (own semantics, PARITY UNPINNED -- the reference only has stubs): the GPU chunk chain against our sequential
    oracle, with the side column and with the tick rule, small chunks (many fix-point rounds / walkers) and the default."""
    from finmlkit_b200 import core
    from finmlkit_b200.bar.logic import _imbalance_bar_indexer, _run_bar_indexer
    ts, px, qty, side = _stream(400_000, seed=9)
    side[::53] = 0
    tr = core.DeviceTrades.upload(ts, px, qty, side, ctx=ctx)
    ref = oracle.imbalance_bar_indexer(side, thr, kind)
    # exact chunk start states from the backward maps (imbalance bars, the default) and the speculative chain (run bars; forced
    # for imbalance bars with FMK_IMBALANCE_NO_EXACT), lane kernel only / walkers only, three chunk sizes
    for ch in (None, "64", "4096"):
        for walk, no_exact in (("0", None), ("1000000000", "1"), ("0", "1")):
            for var, val in (("FMK_CUSUM_CH", ch), ("FMK_CUSUM_WALK_BELOW", walk), ("FMK_IMBALANCE_NO_EXACT", no_exact)):
                if val is None:
                    monkeypatch.delenv(var, raising=False)
                else:
                    monkeypatch.setenv(var, val)
            got = core.imbalance_bar_index(tr, thr, use_side=True, kind=kind).download()
            assert_exact(got[1], ref, f"kind {kind} thr {thr} CH {ch} walk {walk} no_exact {no_exact}")
            assert np.array_equal(got[0], ts[ref])
    for var in ("FMK_CUSUM_CH", "FMK_CUSUM_WALK_BELOW", "FMK_IMBALANCE_NO_EXACT"):
        monkeypatch.delenv(var, raising=False)
    # tick rule (the stub's signature has no side argument): b_t from the prices
    tick_sides = oracle.comp_trade_side_vector(px)
    f = _imbalance_bar_indexer if kind == 0 else _run_bar_indexer
    assert_exact(f(ts, px, qty, thr, ctx=ctx), oracle.imbalance_bar_indexer(tick_sides, thr, kind), "tick rule")
    assert_exact(f(ts, px, qty, thr, sides=side, ctx=ctx), ref, "explicit sides")


def test_imbalance_bars_adversarial_never_coalescing(ctx, monkeypatch):
    """all buys: trajectories from different start states never coalesce (phase of the reset depends on the start), the
    worst case of the chunk chain -- rounds = number of chunks -- must still give the sequential answer."""
    from finmlkit_b200 import core
    n = 60_000
    ts = np.arange(n, dtype=np.int64) * 1000
    side = np.ones(n, np.int8)
    tr = core.DeviceTrades.upload(ts, np.full(n, 100.0), np.ones(n), side, ctx=ctx)
    monkeypatch.setenv("FMK_CUSUM_CH", "96")
    for walk, no_exact in (("0", "1"), ("1000000000", "1"), ("0", None)):
        monkeypatch.setenv("FMK_CUSUM_WALK_BELOW", walk)
        if no_exact:
            monkeypatch.setenv("FMK_IMBALANCE_NO_EXACT", no_exact)
        else:
            monkeypatch.delenv("FMK_IMBALANCE_NO_EXACT", raising=False)
        got = core.imbalance_bar_index(tr, 7.0, use_side=True, kind=0).download()[1]
        assert_exact(got, oracle.imbalance_bar_indexer(side, 7.0, 0), f"all buys walk {walk} no_exact {no_exact}")
        if not no_exact:
            assert ctx.index_stats()["chain_passes"] == 0            # no repair rounds at all: the start states were exact
    for thr in (1.0, 0.5, 1500.0, 2048.0, 3000.0):                      # m = 1, the map-size limit, and beyond it (chain)
        got = core.imbalance_bar_index(tr, thr, use_side=True, kind=0).download()[1]
        assert_exact(got, oracle.imbalance_bar_indexer(side, thr, 0), f"all buys thr {thr}")
    for var in ("FMK_CUSUM_CH", "FMK_CUSUM_WALK_BELOW", "FMK_IMBALANCE_NO_EXACT"):
        monkeypatch.delenv(var, raising=False)


def test_imbalance_bar_kit(ctx):
    from finmlkit_b200.bar.data_model import TradesData
    from finmlkit_b200.bar.kit import ImbalanceBarKit, RunBarKit
    ts, px, qty, side = _stream(100_000, seed=2)
    td = TradesData(ts, px, qty, side=side)
    kit = ImbalanceBarKit(td, 30.0, ctx=ctx)
    df = kit.build_ohlcv()
    ref = oracle.imbalance_bar_indexer(side, 30.0, 0)
    assert np.array_equal(kit.bar_close_indices, ref[1:]) and len(df) == len(ref) - 1
    check_ohlcv([df[c].values for c in ("open", "high", "low", "close", "volume", "vwap", "trades", "median_trade_size")],
                oracle.comp_bar_ohlcv(px, qty, ref), "imbalance kit")
    rk = RunBarKit(td, 200.0, use_side=False, ctx=ctx)
    assert np.array_equal(rk.bar_close_indices, oracle.imbalance_bar_indexer(oracle.comp_trade_side_vector(px), 200.0, 1)[1:])
