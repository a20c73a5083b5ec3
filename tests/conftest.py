import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: CPU-heavy oracle replay")


@pytest.fixture(scope="session")
def engine():
    """One CUDA engine for the whole session (GPU tests only)."""
    import sboxgates_b200 as sb
    eng = sb.LutEngine(0)
    yield eng
    eng.close()
