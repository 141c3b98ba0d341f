"""CPU: the host/device-shared core of the exact dollar-bar algorithm (finmlkit_b200/csrc/dollar_core.h) emulated on the
CPU (tests/cpu/dollar_harness.cpp: same tasks, same certification chain, same serial repair) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from finmlkit_b200.synth import synth_trades

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dollar") / "libdollar_harness.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(HERE, "cpu", "dollar_harness.cpp")])
    lib = C.CDLL(so)
    lib.dollar_emulate.restype = C.c_int64

    def emu(p, v, T, CH):
        p, v = np.ascontiguousarray(p, np.float64), np.ascontiguousarray(v, np.float64)
        out = np.zeros(len(p) + 2, np.int64)
        stats = np.zeros(4, np.int64)
        m = lib.dollar_emulate(p.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_int64(len(p)), C.c_double(T),
                               C.c_int64(CH), out.ctypes.data_as(C.c_void_p), C.c_int64(len(out)), stats.ctypes.data_as(C.c_void_p))
        assert m >= 1
        return out[:m], stats
    return emu


def _build(tmp, name, extra):
    so = str(tmp / name)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", *extra, "-o", so,
                           os.path.join(HERE, "cpu", "dollar_harness.cpp")])
    lib = C.CDLL(so)
    lib.dollar_task_records.restype = C.c_int64

    def recs(p, v, T, CH):
        p, v = np.ascontiguousarray(p, np.float64), np.ascontiguousarray(v, np.float64)
        nt = (len(p) + CH - 1) // CH
        r = np.zeros((nt, 10), np.int64)
        m = np.zeros((nt, 4), np.float64)
        got = lib.dollar_task_records(p.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_int64(len(p)), C.c_double(T),
                                      C.c_int64(CH), r.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p))
        assert got == nt
        return r, m
    return recs


@pytest.mark.parametrize("T,CH", [(1e6, 2048), (1e5, 512), (1048575.0, 512), (float(2 ** 20), 512), (3e4, 257), (1040000.0, 300)])
def test_virtual_chains_equal_explicit_chains(tmp_path_factory, stream, T, CH):
    """The shipped core runs ONE float chain plus integer offsets for the other three residues; the reference build
    (-DDC_EXPLICIT_CHAINS) runs all four chains explicitly.  Wherever the explicit chain rho is usable, the virtual
    chain must report the same end state."""
    tmp = tmp_path_factory.mktemp("xchk")
    virt = _build(tmp, "virt.so", [])
    expl = _build(tmp, "expl.so", ["-DDC_EXPLICIT_CHAINS"])
    _, p, v, _ = stream
    rv, mv = virt(p, v, T, CH)
    re_, me = expl(p, v, T, CH)
    assert np.array_equal(rv[:, :4], re_[:, :4])          # same starts / ends / counts / start state
    checked = 0
    for k in range(1, len(rv)):
        if rv[k, 0] < 0 or rv[k, 1] < 0:
            continue
        for rho in range(4):
            if (re_[k, 8] >> rho) & 1 or (rv[k, 8] >> rho) & 1:
                continue
            assert rv[k, 4 + rho] == re_[k, 4 + rho], (k, rho, rv[k], re_[k])
            checked += 1
    assert checked > 0


@pytest.fixture(scope="module")
def stream():
    return synth_trades(400_000, seed=3)


@pytest.mark.parametrize("T", [1e6, 1e5, 12345.6, 1048575.0, float(2 ** 20), 3e4, 999.99])
@pytest.mark.parametrize("CH", [2048, 257])
def test_synthetic_stream(harness, stream, T, CH):
    _, p, v, _ = stream
    got, stats = harness(p, v, T, CH)
    assert np.array_equal(got, oracle.dollar_bar_indexer(p, v, T))


def test_fast_path_is_taken_at_the_headline_threshold(harness, stream):
    _, p, v, _ = stream
    _, stats = harness(p, v, 1e6, 2048)
    assert stats[1] == 0 and stats[2] > 0     # no serial repair; every non-empty task certified


@pytest.mark.parametrize("T", [1000.0, 5000.0, 1024.0, 777.0])
def test_exact_integer_ties(harness, T):
    rng = np.random.default_rng(0)
    p = rng.integers(90, 110, 100_000).astype(np.float64)
    v = rng.integers(1, 20, 100_000).astype(np.float64)
    got, _ = harness(p, v, T, 512)
    assert np.array_equal(got, oracle.dollar_bar_indexer(p, v, T))


@pytest.mark.parametrize("T", [100.0, 250.0, 1000.0])
def test_decimal_quantised_ties(harness, T):
    rng = np.random.default_rng(1)
    p = np.round(rng.integers(1000, 1100, 100_000) * 0.1, 1)
    v = np.round(rng.integers(1, 2000, 100_000) * 0.001, 3)
    got, _ = harness(p, v, T, 512)
    assert np.array_equal(got, oracle.dollar_bar_indexer(p, v, T))


def test_giant_trades_and_edge_sizes(harness, stream):
    _, p, v, _ = stream
    rng = np.random.default_rng(2)
    v4 = v.copy()
    v4[rng.integers(0, len(v4), 100)] *= 5000
    for T in (1e6, 1e5):
        got, _ = harness(p, v4, T, 2048)
        assert np.array_equal(got, oracle.dollar_bar_indexer(p, v4, T))
    for n, T, CH in ((200_000, 5e7, 256), (50_000, 1e12, 256), (5, 50.0, 2), (1, 50.0, 2)):
        got, _ = harness(p[:n], v[:n], T, CH)
        assert np.array_equal(got, oracle.dollar_bar_indexer(p[:n], v[:n], T))


@pytest.mark.parametrize("case,T", [("giant_4k", 150.0), ("adversarial_6k", 250.0), ("synth_20k", 1e5)])
def test_golden_fixtures(harness, case, T):
    g = np.load(os.path.join(HERE, "golden", case + ".npz"))
    got, _ = harness(g["in_px"], g["in_qty"], T, 64)
    assert np.array_equal(got, g["ref_dollar_idx"])
