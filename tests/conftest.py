import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


_CACHE = {}


def load_golden(name):
    from agentbasedmcmc_b200.abmp import read_abmp
    if name not in _CACHE:
        _CACHE[name] = read_abmp(GOLDEN / name)
    return _CACHE[name]
