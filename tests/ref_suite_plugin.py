"""pytest plugin used by tests/test_gpu_ref_suite.py: puts the installed reference (baseline/_ref) on sys.path and rebinds
its hot path to the GPU implementations (finmlkit_b200.dropin) BEFORE the reference's test modules import their names."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "baseline", "_ref")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("FMK_CONSOLE_LOGGER_LEVEL", "ERROR")

from finmlkit_b200 import dropin  # noqa: E402

N_REBOUND = dropin.install(verbose=True)


def pytest_report_header(config):
    return f"finmlkit_b200.dropin: {N_REBOUND} reference bindings rebound to the GPU implementations"
