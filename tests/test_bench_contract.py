"""bench.py contract on the CPU: the reference arm prints exactly ONE JSON line with the keys the driver reads, ranks > 0
print nothing, and the `ours` arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=300)


def test_reference_arm_prints_one_json_line():
    # the C port (declared fallback) keeps this test fast: the Numba arm pays ~30 s of JIT (covered by the next test)
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ticks", "2e5", "--cpu-sample", "2e5", "--cpu-sub-sample", "2e5"],
             env={"FMK_BENCH_FORCE_PORT": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "ticks/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    # every BASELINE config the CPU can run is in the line
    for k in ("config1", "config3", "config4", "time_bars_1min"):
        assert k in d, k
    assert d["config3"]["value"] > 0 and d["config4"]["value"] > 0 and d["config1"]["kernels"]["value"] > 0


def test_reference_arm_times_the_real_reference_when_installed():
    """baseline/_ref (scripts/install_ref.sh) holds the unmodified finmlkit: the arm must then time Numba, not the port."""
    import pytest
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "finmlkit")):
        pytest.skip("baseline/_ref not installed here")
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ticks", "2e5", "--cpu-sample", "2e5", "--no-sub"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.split("\n") if l.strip()][0])
    assert d["cpu_baseline"]["kind"] == "reference" and "Numba" in d["cpu_baseline"]["impl"]
    assert d["e2e"]["value"] == d["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ticks", "2e5", "--cpu-sample", "2e5", "--no-sub"],
             env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_ours_arm_fails_loudly_without_a_gpu():
    import ctypes
    try:
        ctypes.CDLL("libcuda.so.1")
        has_driver = True
    except OSError:
        has_driver = False
    r = _run(["--steps", "1", "--warmup", "1", "--ticks", "1e5", "--no-e2e"], env={"CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode != 0 and r.stdout.strip() == "", (has_driver, r.stdout[:200])
    assert "no CUDA device" in r.stderr or "CUDA" in r.stderr
