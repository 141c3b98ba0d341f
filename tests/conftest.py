import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no libfmk.so (built artefacts are git-ignored): build it once so that the suite is self-sufficient
    # (nvcc cross-compiles without a GPU); an existing library is left alone -- __graft_entry__.build() owns rebuilds
    from finmlkit_b200 import build as _build
    if not os.path.exists(_build.SO):
        _build.build()


@pytest.fixture(scope="session")
def ctx():
    from finmlkit_b200.core import default_context
    return default_context(0)
